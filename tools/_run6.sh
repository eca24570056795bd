set -x
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_lpips_gpu.py tests/test_vae_gpu.py tests/test_unet_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_tests6.log
tail -12 gpurun_out/r02_tests6.log
timeout 600 python tools/bench_gemm_insitu.py 30 | tee gpurun_out/r02_gemm_insitu_tmares.txt
FD_NO_TMA_RES=1 timeout 600 python tools/bench_gemm_insitu.py 30 | tee gpurun_out/r02_gemm_insitu_notmares.txt
timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -1
FD_NO_TMA_RES=1 timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -1
