set -x
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_sdxl_parity_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/bench_gemm_insitu.py 30 | tee gpurun_out/r02_gemm_insitu_diet.txt
timeout 600 python tools/bench_unet_graph.py 8 2>&1 | tail -1
timeout 900 python bench.py --config sd15 --steps 4 --warmup 3 > gpurun_out/r02_bench_sd15.json 2> gpurun_out/r02_bench_sd15.err; tail -2 gpurun_out/r02_bench_sd15.err; head -c 300 gpurun_out/r02_bench_sd15.json; echo
timeout 900 python bench.py --config pixart --steps 4 --warmup 3 > gpurun_out/r02_bench_pixart.json 2> gpurun_out/r02_bench_pixart.err; tail -2 gpurun_out/r02_bench_pixart.err; head -c 300 gpurun_out/r02_bench_pixart.json; echo
timeout 900 python bench.py --config sd3 --steps 4 --warmup 3 > gpurun_out/r02_bench_sd3.json 2> gpurun_out/r02_bench_sd3.err; tail -2 gpurun_out/r02_bench_sd3.err; head -c 300 gpurun_out/r02_bench_sd3.json; echo
for b in sdxl sd15 pixart sd3; do timeout 900 python bench.py --config sample --backbone $b --steps 3 --warmup 3 > gpurun_out/r02_sample_$b.json 2> gpurun_out/r02_sample_$b.err; tail -1 gpurun_out/r02_sample_$b.err; head -c 200 gpurun_out/r02_sample_$b.json; echo; done
