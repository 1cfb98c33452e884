#!/usr/bin/env bash
# Build libbfilt.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libbfilt.so"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -Xcompiler -O2 -shared \
     --expt-relaxed-constexpr --split-compile 0 "$@" \
     -o "$out" "$here/bfilt_api.cu"
echo "built $out"
