#!/bin/bash
# GPU call 14 (1 GPU): the whole -m gpu suite, then smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
start=$(date +%s)
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/r02_gputest_tail.txt
echo "pytest wall: $(( $(date +%s) - start )) s"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
