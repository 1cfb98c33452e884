#!/usr/bin/env bash
# Build libbfilt.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
#
# The model-templated kernels are compiled as slices of the model registry (tu_kernels.cu, -DBF_TU_KIND / BF_NGROUPS /
# BF_GROUP), one nvcc process per slice, BF_JOBS processes at a time; bfilt_api.cu (host code + the non-template
# kernels) reaches them through the address table of ktable.h.
# (no --split-compile: it makes ptxas register allocation of the big kernels vary from build to build -- measured
#  122 vs 558 bytes of spills for the same source -- and with it the throughput by 15 %)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libbfilt.so"
obj="$here/_build"
mkdir -p "$obj"
# BF_ONLY="fused_1 api": recompile only these objects and relink with the others of the last full build (development)
ONLY=${BF_ONLY:-}
if [[ -z "$ONLY" ]]; then rm -f "$obj"/*.o "$obj"/*.log "$obj"/*.ok; fi

NG_FUSED=${BF_NG_FUSED:-8}; NG_RESIDENT=${BF_NG_RESIDENT:-4}; NG_GSF=${BF_NG_GSF:-8}; NG_MISC=${BF_NG_MISC:-1}
JOBS=${BF_JOBS:-$(nproc)}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2
       --expt-relaxed-constexpr "$@")

# job list: "<object> <source> <defines...>", heaviest first
jobs=()
# The particle-filter translation units are compiled with -fmad=false: every fused multiply-add in them is one the source
# spells out (fmaf), so the general / lean / sharded instantiations of a kernel cannot contract a product-sum differently
# and one filter gives the same bits whichever of them ran it (tests: lean == general, sharded == one GPU).
PF="-fmad=false -DBF_DYNAMIC_CHUNKS"
for ((g = 0; g < NG_FUSED; ++g)); do jobs+=("fused_$g tu_kernels.cu $PF -DBF_TU_KIND=1 -DBF_NGROUPS=$NG_FUSED -DBF_GROUP=$g"); done
jobs+=("large tu_kernels.cu -DBF_TU_KIND=5")
for ((g = 0; g < NG_GSF; ++g)); do jobs+=("gsf_$g tu_kernels.cu -DBF_TU_KIND=3 -DBF_NGROUPS=$NG_GSF -DBF_GROUP=$g"); done
for ((g = 0; g < NG_RESIDENT; ++g)); do jobs+=("resident_$g tu_kernels.cu $PF -DBF_TU_KIND=2 -DBF_NGROUPS=$NG_RESIDENT -DBF_GROUP=$g"); done
for ((g = 0; g < NG_MISC; ++g)); do jobs+=("misc_$g tu_kernels.cu $PF -DBF_TU_KIND=4 -DBF_NGROUPS=$NG_MISC -DBF_GROUP=$g"); done
jobs+=("api bfilt_api.cu $PF -DBF_NG_FUSED=$NG_FUSED -DBF_NG_RESIDENT=$NG_RESIDENT -DBF_NG_GSF=$NG_GSF -DBF_NG_MISC=$NG_MISC")

compile_one() {   # $1 = job line
  set -- $1
  local name=$1 src=$2; shift 2
  if ! nvcc "${FLAGS[@]}" "$@" -c -o "$obj/$name.o" "$here/$src" > "$obj/$name.log" 2>&1; then
    echo "FAILED: $name" >&2; cat "$obj/$name.log" >&2; return 1
  fi
  cat "$obj/$name.log"
  : > "$obj/$name.ok"
}

for job in "${jobs[@]}"; do
  name=${job%% *}
  if [[ -n "$ONLY" && " $ONLY " != *" $name "* ]]; then continue; fi
  rm -f "$obj/$name.ok"
  while (( $(jobs -rp | wc -l) >= JOBS )); do wait -n || true; done
  compile_one "$job" &
done
wait || true
for job in "${jobs[@]}"; do
  name=${job%% *}
  [[ -f "$obj/$name.ok" ]] || { echo "build failed ($name)" >&2; exit 1; }
done

nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$out" "$obj"/*.o
echo "built $out"
