#!/bin/bash
# GPU call 16 (2 GPUs): the 2-rank NCCL equivalence test (full output kept)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -x -q -m gpu > gpurun_out/r02_dp_nccl_2gpu_full.txt 2>&1
tail -30 gpurun_out/r02_dp_nccl_2gpu_full.txt
