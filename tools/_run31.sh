#!/bin/bash
# GPU call 31: GPU test files not re-run since the host-side changes of this session (schedulers on the host under any
# default device, LinearPack K padding, collation / config semantics): goldens, LPIPS, reference-golden CUDA replays
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 170 python -m pytest tests/test_reference_golden.py tests/test_golden.py tests/test_lpips_gpu.py tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | grep -v Warning | tail -12 ) | tee gpurun_out/r02_regression2_gpu.txt
