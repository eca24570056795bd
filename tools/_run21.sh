#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FD_ATTN_DIAG=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd1 -s 2 -c 1 -o gpurun_out/r02_prof_attn_diag3 -f python tools/bench_attn.py > gpurun_out/r02_ncu_attn_diag3.log 2>&1; tail -2 gpurun_out/r02_ncu_attn_diag3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd1 -s 2 -c 1 -o gpurun_out/r02_prof_attn_now -f python tools/bench_attn.py > gpurun_out/r02_ncu_attn_now.log 2>&1; tail -2 gpurun_out/r02_ncu_attn_now.log
