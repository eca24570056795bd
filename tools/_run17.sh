#!/bin/bash
# GPU call 17 (1 GPU): attention early-TMA A/B (+ its tests), tile-group size A/B on the in-situ GEMMs, bench lines of the
# other configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_attention_gpu.py -x -q -m gpu -k "attn or attention" 2>&1 | tail -3
for i in 1 2; do
FD_ATTN_EARLY=0 timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -3 | sed 's/^/EARLY=0: /'
FD_ATTN_EARLY=1 timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -3 | sed 's/^/EARLY=1: /'
done
for gm in default 15 5 4; do
  if [ "$gm" = default ]; then unset FD_GROUP_M; else export FD_GROUP_M=$gm; fi
  timeout 300 python tools/bench_gemm_insitu.py 30 ff2,o,ff1,qkv,o640,ff1b 2>&1 | grep TF | sed "s/^/GROUP_M=$gm: /"
done
unset FD_GROUP_M
for cfg in sd15 pixart sd3; do
  timeout 900 python bench.py --config $cfg --steps 4 --warmup 3 2> gpurun_out/r02_bench_$cfg.err | grep '^{' > gpurun_out/r02_bench_$cfg.json
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_$cfg.json').read()); print('$cfg', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
