#!/bin/bash
# Build libflashb200.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
mkdir -p ../lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static"
OBJS=""
pids=()
for f in fd_api fd_gemm fd_norm fd_elem fd_attn fd_attn_bwd fd_attn_generic fd_attn_bwd_generic; do
  [ -f $f.cu ] || continue
  if [ ! -f ../lib/$f.o ] || [ $f.cu -nt ../lib/$f.o ] || [ fd_common.cuh -nt ../lib/$f.o ] || [ fd_host.h -nt ../lib/$f.o ] || [ ../../include/flashb200.h -nt ../lib/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o ../lib/$f.o &
    pids+=($!)
  fi
  OBJS="$OBJS ../lib/$f.o"
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../lib/libflashb200.so $OBJS
echo "built $(realpath ../lib/libflashb200.so)"
