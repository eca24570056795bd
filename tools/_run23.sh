#!/bin/bash
# GPU call 23 (1 GPU): the whole -m gpu suite + smoke + default bench with the final kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r02_gputest_tail.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_final.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['teacher_evaluation'], d['roofline']['attention_fwd'], d['clocks'])"
