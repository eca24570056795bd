"""2-process NCCL test of the data-parallel CUDA path: two ranks x batch 2 on the hand-written kernels,
gradients averaged by the flat-bucket NCCL all-reduce, end with the same LoRA / discriminator update as ONE rank with
batch 4 fed the same samples and draws.  Needs 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_dp_nccl_gpu.py`
(skipped on a 1-GPU box; the log of the 2-GPU run is committed under profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_ranks_nccl_equal_one_rank_double_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    sys.path.insert(0, HERE)
    import dp_nccl_worker as W
    port = 29600 + os.getpid() % 300
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "dp_nccl_worker.py"),
                    str(tmp_path)], check=True, timeout=900)
    dp = torch.load(os.path.join(tmp_path, "dp.pt"))
    dev = torch.device("cuda", 0)
    model, pipe = W.build(dev, seed=1234)
    init = torch.load(os.path.join(tmp_path, "init.pt"))  # rank 0's weights: what the broadcast made every replica
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n in init:
                p.copy_(init[n].to(dev))
    before = {n: p.detach().cpu().clone() for n, p in model.named_parameters()}
    batch, draws = W.data(4, dev)
    pipe.training_step(batch, 0, draws=draws)
    torch.cuda.synchronize()
    moved, worst = 0, 1.0
    for n, p in model.named_parameters():
        if n in dp:
            d_ref = p.detach().cpu() - before[n]
            d_dp = dp[n] - before[n]
            if d_ref.abs().max() > 0:
                moved += 1
                c = float(torch.dot(d_ref.reshape(-1), d_dp.reshape(-1)) / (d_ref.norm() * d_dp.norm() + 1e-30))
                worst = min(worst, c)
                # same kernels and per-sample math, but M = 2 x 1024 rows per rank instead of 4 x 1024 selects other GEMM
                # tiles / K splits (other fp32 summation orders before the bf16 rounding of every activation): measured
                # 2.1e-2 on the noisiest LoRA-B tensor at cos 0.9998 (profiles/r02_dp_nccl_2gpu.txt); the fp32 gloo twin of
                # this test (tests/test_dp_gloo_cpu.py) holds 2e-3
                assert (d_dp - d_ref).norm() <= 5e-2 * d_ref.norm() + 1e-9, (n, c)
    assert moved > 10 and worst > 0.999, (moved, worst)
