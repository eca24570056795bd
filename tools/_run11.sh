#!/bin/bash
# GPU call 11 (2 GPUs): NCCL equivalence test with full output + 2-GPU bench after the rank-0-only-step fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -x -q -m gpu > gpurun_out/r02_dp_nccl_2gpu_full.txt 2>&1
tail -80 gpurun_out/r02_dp_nccl_2gpu_full.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
echo "bench rc=$?"
tail -5 gpurun_out/r02_bench_2gpu.err
head -c 300 gpurun_out/r02_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['allreduce'], d['n_gpus'])"
