set -x
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_tests3a.log
tail -5 gpurun_out/r02_tests3a.log
timeout 300 python tools/diag_attn_bwd.py 2>&1 | tail -6 | tee gpurun_out/r02_diag_attn_bwd.txt
FD_ATTN_V=1 timeout 300 python tools/bench_attn.py 2>&1 | grep attn | tee gpurun_out/r02_bench_attn_v1.txt
FD_ATTN_V=2 timeout 300 python tools/bench_attn.py 2>&1 | grep -E "attn|groupnorm" | tee gpurun_out/r02_bench_attn_v2.txt
timeout 600 python tools/bench_gemm.py 0 2>&1 | tee gpurun_out/r02_bench_gemm_auto.txt
FD_TAIL_SPLIT=0 FD_SK_MIN=99999 timeout 600 python tools/bench_gemm.py 0 2>&1 | tee gpurun_out/r02_bench_gemm_plain.txt
FD_SK_MIN=0 timeout 600 python tools/bench_gemm.py 0 2>&1 | tee gpurun_out/r02_bench_gemm_sk0.txt
FD_ATTN_V=1 timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -1
FD_ATTN_V=2 timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -1
FD_ATTN_V=2 FD_TAIL_SPLIT=0 FD_SK_MIN=99999 timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -1
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_backward_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_tests3b.log
tail -5 gpurun_out/r02_tests3b.log
timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -3 gpurun_out/r02_bench_a.err; head -c 1500 gpurun_out/r02_bench_a.json
