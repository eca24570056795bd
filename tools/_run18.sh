#!/bin/bash
# GPU call 18: attention timing diagnostics (results intentionally wrong under FD_ATTN_DIAG) + sample-config lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 0 1 2 3; do
FD_ATTN_DIAG=$d timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -2 | sed "s/^/DIAG=$d: /"
done
for bb in sdxl sd15 pixart sd3; do
  timeout 600 python bench.py --config sample --backbone $bb 2> gpurun_out/r02_sample_$bb.err | grep '^{' > gpurun_out/r02_sample_$bb.json
  python -c "
import json; d=json.loads(open('gpurun_out/r02_sample_$bb.json').read()); print('$bb', d['value'], d['unit'], [ (r['batch'], round(r['latency_ms'],1)) for r in d.get('rows',[])][:6])"
done
