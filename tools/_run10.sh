set -x
nvidia-smi -L
timeout 1200 python -m pytest tests/test_dp_nccl_gpu.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_dp_nccl_2gpu.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -3 gpurun_out/r02_bench_2gpu.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['allreduce'], d['n_gpus'])"
