#!/bin/bash
# GPU call 32: exactness property of the fused CFG + solver kernel at the BASELINE latent size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 90 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k point_mass 2>&1 | grep -v Warning | tail -15 ) | tee gpurun_out/r02_solver_property_gpu.txt
