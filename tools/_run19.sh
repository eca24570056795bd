#!/bin/bash
# GPU call 19: separate K / V rings in the attention forward — parity tests, isolated timing, graphed teacher evaluation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_dit_gpu.py tests/test_backward_gpu.py -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -6; done
timeout 300 python tools/bench_unet_graph.py 8 2>&1 | tail -1
