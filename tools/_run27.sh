#!/bin/bash
# GPU call 27: new parity tests (odd input widths, the Transformer2DModel variant of the reference's wrapper test, SD3 VAE)
# + regression of the files whose code paths changed (LinearPack K padding is a no-op for widths that are multiples of 8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_unet_gpu.py::test_odd_context_and_vector_widths tests/test_dit_gpu.py::test_reference_test_transformer_variant tests/test_vae_gpu.py -q -m gpu 2>&1 | grep -v Warning | tail -40 ) | tee gpurun_out/r02_new_parity_gpu.txt
( timeout 420 python -m pytest tests/test_dit_gpu.py tests/test_unet_gpu.py tests/test_backward_gpu.py -x -q -m gpu --deselect tests/test_dit_gpu.py::test_reference_test_transformer_variant --deselect tests/test_unet_gpu.py::test_odd_context_and_vector_widths 2>&1 | grep -v Warning | tail -15 ) | tee gpurun_out/r02_regression_gpu.txt
