"""torchrun worker of tests/test_dp_nccl_gpu.py: every rank builds the product model ON ITS OWN GPU with ITS OWN seed
(the start-up broadcast must make the replicas rank 0's), runs one CUDA training step on its half of the batch with
the NCCL flat-bucket all-reduce, and rank 0 saves the trained parameters."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "flash-diffusion_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def build(dev, seed):
    from flash import recipes
    from flash.trainer import TrainingConfig, TrainingPipeline
    model, _ = recipes.build_tiny_distillation(dev, seed=seed, lora_b_std=0.02)
    pipe = TrainingPipeline(model, TrainingConfig(optimizers_name=["SGD", "SGD"], learning_rates=[0.05, 0.05],
                                                  trainable_params=[["student_denoiser"], ["discriminator."]]))
    pipe.configure_optimizers()
    return model, pipe


def data(B, dev):
    from flash import recipes
    batch = recipes.synthetic_batch(B, 32, 77, 96, 48, seed=7, device=dev, image_px=256.0)
    g = torch.Generator(device=dev).manual_seed(11)
    shape = (B, 4, 32, 32)
    draws = dict(noise=torch.randn(shape, device=dev, generator=g), start_idx=1, guidance=5.0,
                 dmd_noise=torch.randn(shape, device=dev, generator=g),
                 dmd_timestep=torch.tensor([700, 120, 333, 901][:B], device=dev), dmd_guidance=4.0,
                 gan_noise=torch.randn(shape, device=dev, generator=g),
                 gan_timesteps=torch.tensor([250, 750, 10, 500][:B], device=dev))
    return batch, draws


def slice_(d, lo, hi):
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] >= hi else v) for k, v in d.items()}


def main():
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model, pipe = build(dev, seed=1234 + 100 * rank)
    keep = lambda n, p: p.requires_grad or "discriminator" in n

    def checksum():
        v = torch.stack([p.detach().double().sum() for n, p in model.named_parameters() if keep(n, p)])
        lo, hi = v.clone(), v.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return float((hi - lo).abs().max())

    # configure_optimizers() ended with TrainingPipeline.sync_replicas(): the per-rank seeds are gone
    assert checksum() == 0.0, "start-up broadcast left the replicas different"
    if rank == 0:
        torch.save({n: p.detach().cpu() for n, p in model.named_parameters() if keep(n, p)},
                   os.path.join(out_dir, "init.pt"))
    batch, draws = data(2 * world, dev)
    pipe.training_step(slice_(batch, 2 * rank, 2 * rank + 2), 0, draws=slice_(draws, 2 * rank, 2 * rank + 2))
    torch.cuda.synchronize()
    assert checksum() == 0.0, "replicas diverged after one all-reduced step"
    if rank == 0:
        torch.save({n: p.detach().cpu() for n, p in model.named_parameters() if keep(n, p)},
                   os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
