#!/bin/bash
# GPU call 30: unconditional UNet2DModel wrapper (parity + the reference's own TestDiffusersUNet2DWrapper, payload appended
# by the caller), then smoke() and a short default bench of the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 240 python -m pytest tests/test_unet2d_gpu.py -q -m gpu 2>&1 | grep -v Warning | tail -40 ) | tee gpurun_out/r02_unet2d_gpu.txt
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee gpurun_out/r02_smoke_final.txt
( timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>gpurun_out/r02_bench_final.err | tail -1 ) > gpurun_out/r02_bench_final.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
print("bench:", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"])
PY
