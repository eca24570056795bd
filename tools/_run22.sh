#!/bin/bash
cd "$(dirname "$0")/.."
FD_ATTN_POLY=-1 timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -2 | sed "s/^/NOEXP: /"
for v in 0 1 2 3; do
FD_ATTN_VAR=$v timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -2 | sed "s/^/VAR=$v: /"
done
FD_ATTN_VAR=3 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q -m gpu -k "attn or attention or unet_forward" 2>&1 | tail -3
