#!/bin/bash
# GPU call 13 (1 GPU): reference-golden CUDA test, colstats unet test, attention barrier A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_golden.py tests/test_unet_gpu.py -x -q -m gpu 2>&1 | tail -15
for i in 1 2; do
timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -3
FD_ATTN_BAR256=1 timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -3 | sed 's/^/BAR256: /'
done
