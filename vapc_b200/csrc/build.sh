#!/usr/bin/env bash
# Build libvapc_b200.so for B200 (sm_100a).  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${1:-$HERE/../libvapc_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS=(keys.cu scan.cu sort.cu bucket.cu cell_sort.cu segment_reduce.cu stats.cu bitemporal.cu mode_mask.cu neighbors.cu laz.cu)
cd "$HERE"
"$NVCC" ${VAPC_NVCC_FLAGS:-} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC,-O3,-fvisibility=hidden -shared -cudart static \
  -Xlinker -soname,libvapc_b200.so "${SRCS[@]}" -o "$OUT"
echo "built $OUT"
