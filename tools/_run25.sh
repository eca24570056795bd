#!/bin/bash
# GPU call 25: the SD3 VAE parity tests + the three full-size reference example scripts, unmodified, 2 training steps each
# (payload = tools/example_run_cmd.py sdxl sd3 pixart, passed as $1 by the caller because /root/reference is not on the box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
free -g | head -2
( timeout 300 python -m pytest tests/test_vae_gpu.py -x -q -m gpu -k "sd3" 2>&1 | tail -5 ) | tee gpurun_out/r02_sd3_vae_gpu.txt
