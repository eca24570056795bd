#!/bin/bash
# GPU call 15 (1 GPU): default bench (wall time), reference arm, CUPTI step profile, ncu traffic capture of the
# dominant in-situ GEMMs, ncu launch list of the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench default rc=$? wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/r02_bench_default.err
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
echo "bench reference rc=$? wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/r02_bench_reference.err; head -c 600 gpurun_out/r02_bench_reference.json; echo
timeout 600 python tools/profile_step.py 16 > gpurun_out/r02_step_profile.txt 2>&1; head -12 gpurun_out/r02_step_profile.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 12 -c 8 -o gpurun_out/r02_prof_gemm_insitu2 -f python tools/bench_gemm_insitu.py 1 ff1,ff2,qkv,o > gpurun_out/r02_ncu_gemm2.log 2>&1; tail -3 gpurun_out/r02_ncu_gemm2.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 14000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_launches_bench.log 2>&1; tail -2 gpurun_out/r02_launches_bench.log; wc -l gpurun_out/r02_launches_bench.csv
