set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
lscpu | head -20
python -m pytest tests/test_sdxl_parity_gpu.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_sdxl_parity.log
python tools/bench_eager.py > gpurun_out/r02_eager.json 2> gpurun_out/r02_eager.err
tail -3 gpurun_out/r02_eager.err
cat gpurun_out/r02_sdxl_parity.log | tail -15
