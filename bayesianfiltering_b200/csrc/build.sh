#!/usr/bin/env bash
# Build libbfilt.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
# (no --split-compile: it makes ptxas register allocation of the big kernels vary from build to build -- measured
#  122 vs 558 bytes of spills for the same source -- and with it the throughput by 15 %)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libbfilt.so"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -Xcompiler -O2 -shared \
     --expt-relaxed-constexpr "$@" \
     -o "$out" "$here/bfilt_api.cu"
echo "built $out"
