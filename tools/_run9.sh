set -x
timeout 1500 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
tail -2 gpurun_out/r02_bench_c.err; head -c 400 gpurun_out/r02_bench_c.json
