#!/bin/bash
cd "$(dirname "$0")/.."
for sp in 0 1; do
FD_ATTN_SPIN=$sp timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -3 | sed "s/^/SPIN=$sp: /"
done
FD_ATTN_SPIN=1 FD_ATTN_DIAG=3 timeout 300 python tools/bench_attn.py 2>&1 | grep "^attn" | head -2 | sed "s/^/SPIN=1 DIAG=3: /"
FD_ATTN_SPIN=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attn or attention" 2>&1 | tail -2
