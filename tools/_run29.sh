#!/bin/bash
# GPU call 29: where the vector-conditioned Transformer2D variant goes non-finite (diagnostic), the new parity tests, the
# reference's own wrapper tests (payload appended by the caller: tools/reference_tests_cmd.py wrapper)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 150 python tools/diag_nan.py 2>&1 | grep -v Warning | tail -150 | cut -c1-230 ) > gpurun_out/r02_diag_nan.txt; grep -n "FIRST\|=====\|ref absmax" gpurun_out/r02_diag_nan.txt
( timeout 200 python -m pytest tests/test_unet_gpu.py::test_odd_context_and_vector_widths tests/test_dit_gpu.py::test_reference_test_transformer_variant tests/test_dit_gpu.py::test_small_pixart_forward tests/test_vae_gpu.py -q -m gpu 2>&1 | grep -v Warning | tail -30 ) | tee gpurun_out/r02_new_parity_gpu.txt
