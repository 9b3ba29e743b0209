#!/usr/bin/env bash
# Build libegnn_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC,-O3,-Wall -shared ${EGNN_NVCC_EXTRA:-} \
    -o "$out/libegnn_b200.so" "$here/egnn_cabi.cu"
echo "built $out/libegnn_b200.so"
