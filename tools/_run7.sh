set -x
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_gpu_tests_full.log
tail -8 gpurun_out/r02_gpu_tests_full.log
timeout 1500 python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
tail -2 gpurun_out/r02_bench_b.err; head -c 600 gpurun_out/r02_bench_b.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
tail -2 gpurun_out/r02_bench_ref.err; head -c 400 gpurun_out/r02_bench_ref.json
