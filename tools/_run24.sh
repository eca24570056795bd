#!/bin/bash
# GPU call 24 (4 GPUs): scaling sanity of the final code (overlapped all-reduce + CUDA graphs + NCCL at 4 ranks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
echo "rc=$?"; tail -3 gpurun_out/r02_bench_4gpu.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_4gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['allreduce'])"
