set -x
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_gemm_tests.log
tail -5 gpurun_out/r02_gemm_tests.log
timeout 600 python tools/bench_gemm.py 0 > gpurun_out/r02_bench_gemm_sk.txt 2>&1
FD_NO_STREAMK=1 FD_NO_TMA_STORE=1 timeout 600 python tools/bench_gemm.py 0 > gpurun_out/r02_bench_gemm_nosk.txt 2>&1
FD_NO_TMA_STORE=1 timeout 600 python tools/bench_gemm.py 0 > gpurun_out/r02_bench_gemm_sk_notma.txt 2>&1
cat gpurun_out/r02_bench_gemm_sk.txt gpurun_out/r02_bench_gemm_nosk.txt gpurun_out/r02_bench_gemm_sk_notma.txt
timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -2
FD_NO_STREAMK=1 FD_NO_TMA_STORE=1 timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -2
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_backward_gpu.py tests/test_sdxl_parity_gpu.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_tests2.log
tail -25 gpurun_out/r02_tests2.log
