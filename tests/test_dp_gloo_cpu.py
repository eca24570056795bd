"""World-size-2 gloo test of the data-parallel path: two ranks with batch B each, after the
flat-bucket all-reduce, end up with the same LoRA / discriminator update as one rank with batch 2B fed the same
samples and draws.  Denoisers are the fp32 oracle UNets (host-logic test; the product denoiser is CUDA-only)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(seed=0):
    for p in (ROOT, os.path.join(ROOT, "flash-diffusion_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from test_flash_step_cpu import _model
    from flash.trainer import TrainingConfig, TrainingPipeline
    model = _model(seed=seed)
    pipe = TrainingPipeline(model, TrainingConfig(optimizers_name=["SGD", "SGD"], learning_rates=[0.1, 0.1],
                                                  trainable_params=[["student_denoiser"], ["discriminator."]]))
    pipe.configure_optimizers()
    return model, pipe


def _data(B, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_flash_step_cpu import _batch, _draws
    return _batch(B=B, seed=seed), _draws(B=B, seed=seed + 1)


def _slice(d, lo, hi):
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] >= hi else v) for k, v in d.items()}


def _worker(rank, world, port, out_dir, seed_per_rank=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(2)
    # seed_per_rank: every process initialises LoRA A / the discriminator differently, as real launches do; the
    # start-up broadcast in TrainingPipeline.sync_replicas (Lightning DDP's wrap-time sync) must make them rank 0's
    model, pipe = _build(seed=100 * rank if seed_per_rank else 0)
    batch, draws = _data(4, 5)
    b = _slice(batch, 2 * rank, 2 * rank + 2)
    d = _slice(draws, 2 * rank, 2 * rank + 2)
    pipe.training_step(b, 0, draws=d)
    if rank == 0:
        torch.save({n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad or "discriminator" in n},
                   os.path.join(out_dir, "dp.pt"))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("seed_per_rank", [False, True])
def test_two_ranks_equal_one_rank_double_batch(tmp_path, seed_per_rank):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    port = 29500 + (os.getpid() + 7 * int(seed_per_rank)) % 500
    mp.spawn(_worker, args=(2, port, str(tmp_path), seed_per_rank), nprocs=2, join=True)
    dp = torch.load(os.path.join(tmp_path, "dp.pt"))
    torch.set_num_threads(4)
    model, pipe = _build()
    batch, draws = _data(4, 5)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    pipe.training_step(batch, 0, draws=draws)
    moved = 0
    for n, p in model.named_parameters():
        if n in dp:
            delta_ref = p.detach() - before[n]
            delta_dp = dp[n] - before[n]
            if delta_ref.abs().max() > 0:
                moved += 1
                # batch-mean losses: mean over 2 ranks of mean over B == mean over 2B, except the DMD weight
                # (per-sample mean-abs normaliser, unaffected) -> identical up to fp32 reduction order
                assert torch.allclose(delta_dp, delta_ref, rtol=2e-3, atol=1e-6), n
    assert moved > 10
