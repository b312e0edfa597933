#!/bin/bash
# Builds dang_b200/libdang_gpu.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../libdang_gpu.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-O2,-Wall -shared -o "$out" "$here/dang_gpu.cu" -ldl "$@"
echo "built $out"
