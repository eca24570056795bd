set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 9 -c 6 -o gpurun_out/r02_prof_gemm_insitu -f python tools/bench_gemm_insitu.py 1 o640,o,ff1b > gpurun_out/r02_ncu_gemm.log 2>&1
tail -3 gpurun_out/r02_ncu_gemm.log
timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_unet_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -1
