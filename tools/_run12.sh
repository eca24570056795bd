#!/bin/bash
# GPU call 12 (1 GPU): GroupNorm statistics from the producer epilogue (colstats) — parity tests, A/B on the graphed
# teacher evaluation, attention forward after the per-quarter barrier change
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q -m gpu -k "colstats or groupnorm or unet or kv_cache or producer" 2>&1 | tail -15
for i in 1 2; do
timeout 300 python tools/bench_unet_graph.py 8 2>&1 | tail -1
FD_NO_COLSTATS=1 timeout 300 python tools/bench_unet_graph.py 8 2>&1 | tail -1 | sed 's/^/NO_COLSTATS: /'
done
timeout 300 python tools/bench_attn.py 2>&1 | tail -12
