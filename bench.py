#!/usr/bin/env python
"""bench.py — Flash-Diffusion distillation-step throughput on B200 (BASELINE.json metric, config 2).

  python bench.py --gpus N --steps K --warmup W            our arm   (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the oracle (fp32 PyTorch restatement of
                                                           the reference's diffusers path; the reference itself is not
                                                           installable here) on the host CPU cores

Workload (config.workload): SDXL UNet 1024x1024 (latent 128x128) LoRA-rank-64 distillation step, batch 4 per GPU,
bf16, K=32 trailing DPM-Solver++ teacher, DMD + lsgan, l2 distill, synthetic latents / text embeddings, random-init
weights (seed 1234).  A "step" is one full `TrainingPipeline.training_step` (both optimizer turns, reference
src/flash/trainer/trainer.py:169-218).  The teacher-rollout length depends on the sampled start index
(flash_diffusion_model.py:167,289); the timed steps pin start_idx to 0, 8, 16, 24 in turn (the four mixture modes,
uniform weights = stage 3 of flash_sdxl.yaml:27-32, E[n] = 20) so every run does the same work.

One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "distillation images/sec"
UNIT = "images/s"
BATCH_PER_GPU = 4
MODES = [0, 8, 16, 24]                      # start indices of the 4 mixture modes, K = 32
F_FWD, F_DM = 6.76e12, 2.93e12              # SDXL UNet FLOPs / sample: full forward, down+mid only
F_LORA_DW = 2 * 2 * 560 * 0                 # (accounted inside the measured kernels; negligible: see DESIGN.md)


def flops_per_image(n):
    """2*[(4+2n)F + 2 F_dm] + F (student dX) + F_dm (GAN dX through the teacher)."""
    return 2 * ((4 + 2 * n) * F_FWD + 2 * F_DM) + F_FWD + F_DM


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
# Bounded CPU sample (a full SDXL step is ~2.5 PFLOP per batch: hours on CPU): the SDXL UNet MID BLOCK
# (ResnetBlock2D + Transformer2D depth 10 at 1024 tokens x 1280 channels, 20 heads, ctx 77x2048 + ResnetBlock2D;
# examples/train_flash_sdxl.py:66-118) of the ORACLE (fp32 PyTorch restatement of the reference's diffusers path) at
# batch 1.  Its analytic cost is F_MID FLOPs; the measured CPU FLOP rate is extrapolated to the whole step with
# the reference step structure (SURVEY.md §8d formula).
F_MID = 797.28e9
_CPU = {}


def cpu_threads():
    return min(os.cpu_count() or 1, 32)


def _cpu_mid_block(lora):
    key = "lora" if lora else "plain"
    if key not in _CPU:
        from oracle.unet import LoraConfig, MidBlock, UNet2DConditionOracle  # noqa: F401
        torch.manual_seed(1234)
        mid = MidBlock(1280, 1280, 32, 1e-5, dict(heads=20, dim_head=64, num_layers=10, cross_attention_dim=2048,
                                                   groups=32, use_linear_projection=True))
        if lora:
            UNet2DConditionOracle.add_adapter(mid, LoraConfig(r=64, lora_alpha=64, init_lora_weights="gaussian",
                                                              target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
        else:
            for p_ in mid.parameters():
                p_.requires_grad = False
        _CPU[key] = mid
    return _CPU[key]


def cpu_sample(backward=False):
    """seconds for one mid-block forward (no grad) or forward+backward (LoRA) at batch 1 on the host cores."""
    torch.set_num_threads(cpu_threads())
    mid = _cpu_mid_block(lora=backward)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 1280, 32, 32, generator=g)
    temb = torch.randn(1, 1280, generator=g)
    ctx = torch.randn(1, 77, 2048, generator=g)
    if backward:
        x.requires_grad_(True)
        t0 = time.perf_counter()
        mid(x, temb, ctx).square().mean().backward()
        return time.perf_counter() - t0
    with torch.no_grad():
        t0 = time.perf_counter()
        mid(x, temb, ctx)
        return time.perf_counter() - t0


def cpu_images_per_sec(t_fwd, t_fwd_bwd=None, n=20):
    """images/s of the reference step: forward FLOPs 2[(4+2n)F + 2F_dm] at the measured forward rate,
    backward FLOPs (F + F_dm) at the measured backward rate (fwd+bwd sample minus a forward)."""
    rate_f = F_MID / t_fwd
    if t_fwd_bwd is not None and t_fwd_bwd > t_fwd:
        rate_b = 2.0 * F_MID / (t_fwd_bwd - t_fwd)        # dX + dW(LoRA) ~ 2x forward FLOPs of the sample
    else:
        rate_b = rate_f
    fwd = 2 * ((4 + 2 * n) * F_FWD + 2 * F_DM)
    bwd = 2 * (F_FWD + F_DM)
    return 1.0 / (fwd / rate_f + bwd / rate_b)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    cpu_sample()                                   # build + first-touch outside the timed samples
    t_fb = cpu_sample(backward=True)
    t_fb = min(t_fb, cpu_sample(backward=True))
    vals, tf = [], []
    for i in range(args.warmup + args.steps):
        t = cpu_sample()
        if i >= args.warmup:
            tf.append(t)
            vals.append(cpu_images_per_sec(t, t_fb))
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * BATCH_PER_GPU / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                             "sample": "per step: one forward of the SDXL UNet mid block (797 GFLOP: Res + Transformer2D "
                                       "depth 10 @1024 tok x 1280 ch + Res) of the fp32 oracle at batch 1; fwd+bwd (LoRA) "
                                       "timed twice up front; images/s extrapolated by FLOPs with the reference step "
                                       "structure at E[n]=20 (616 TFLOP/image)",
                             "t_fwd_s": sum(tf) / len(tf), "t_fwd_bwd_s": t_fb,
                             "cpu_tflops_fwd": F_MID / (sum(tf) / len(tf)) / 1e12, "host_cores": os.cpu_count()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line))


def config_dict(n_gpus):
    return {"workload": "Flash-SDXL UNet 1024x1024 (latent 128x128) LoRA-rank-64 distillation step "
                        "(student fwd+bwd, K=32 DPM-Solver++ teacher CFG rollout, DMD, lsgan GAN), bf16, batch 4/GPU",
            "global_batch": BATCH_PER_GPU * n_gpus, "batch_per_gpu": BATCH_PER_GPU, "latent": [4, 128, 128],
            "context": [77, 2048], "vector": 2816, "lora_rank": 64, "K": 32,
            "start_idx_schedule": MODES, "expected_teacher_steps": 20,
            "parallelism": f"dp{n_gpus}", "l2_policy": "inputs and activations larger than L2 (weights 5.1 GB bf16 per UNet)",
            "weights": "random-init seed 1234", "distill_loss": "l2 (lpips needs offline-unavailable weights)"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    from flash import recipes
    from flash.b200 import lib as fdlib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = fdlib.load()
    lib.fd_launch_count.restype = __import__("ctypes").c_longlong

    model, pipe = recipes.build_sdxl_distillation(dev)
    B = BATCH_PER_GPU

    def host_batch(i):
        return recipes.synthetic_batch(B, 128, 77, 2048, 1280, seed=1234 + rank + 1000 * i, pin=True)

    def draws(i):
        return {"start_idx": MODES[i % len(MODES)]}

    def step(batch, i):
        return pipe.training_step(batch, i, draws=draws(i))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (weight packing, autotuned heuristics, allocator)
    for w in range(args.warmup):
        step({k: v.to(dev) for k, v in host_batch(-1 - w).items()}, 3)     # start_idx 24: shortest rollout
    resident = [{k: v.to(dev) for k, v in host_batch(i).items()} for i in range(args.steps)]

    from flash.b200 import graphs

    def launches_now():
        return lib.fd_launch_count() + graphs.REPLAYED_LAUNCHES

    def timed(fn):
        barrier()
        l0 = launches_now()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), launches_now() - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # (1) kernel/device throughput: inputs already resident in HBM
    ms_dev, launches = timed(lambda: [step(resident[i], i) for i in range(args.steps)])
    # (2) end to end: pinned host batch -> H2D -> step -> loss read back (D2H), every step
    hosts = [host_batch(100 + i) for i in range(args.steps)]
    h2d = sum(v.numel() * v.element_size() for v in hosts[0].values())
    sink = []

    def e2e_loop():
        for i in range(args.steps):
            b = {k: v.to(dev, non_blocking=True) for k, v in hosts[i].items()}
            out = step(b, i)
            sink.append((float(out["loss_optimizer_0"]), float(out["loss_optimizer_1"])))

    ms_e2e, _ = timed(e2e_loop)
    # (3) same steps with the output-preserving dead-work elision on the discriminator turn: reported
    #     separately, never as `value`
    model.elide_unused_generator_pass = True
    ms_lean, _ = timed(lambda: [step(resident[i], i) for i in range(args.steps)])
    model.elide_unused_generator_pass = False
    clocks = sampler.stop() if rank == 0 else None

    imgs = B * world * args.steps
    value = imgs / (ms_dev / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)

    # roofline of the dominant kernel family (tcgen05 GEMM / implicit-GEMM conv): per-launch CUDA-event timing on
    # the launch stream during one extra, untimed teacher CFG evaluation at batch 2B (the op that is ~85% of a step)
    roof = None
    if rank == 0:
        import ctypes
        peak_tf, peak_hbm, peak_src = peaks()
        model.use_cuda_graphs = False            # eager launches so that every launch gets its CUDA events
        lib.fd_profile_enable(1)
        with torch.no_grad():
            b0 = resident[0]
            cond = model.conditioner(b0, set_ucg_rate_zero=True)
            unc = model.conditioner(b0, ucg_keys=model.ucg_keys)
            ts = torch.full((B,), 500, device=dev)
            model._teacher_pair(model.teacher_denoiser, b0["image"], ts, cond, unc)
        lib.fd_profile_enable(0)
        model.use_cuda_graphs = True
        ms = (ctypes.c_double * 4)()
        fl = (ctypes.c_double * 4)()
        cnt = (ctypes.c_longlong * 4)()
        lib.fd_profile_summary(ms, fl, cnt, 4)
        g_ms, g_fl, g_n = ms[0] + ms[1], fl[0] + fl[1], cnt[0] + cnt[1]
        ach = g_fl / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "fd::gemm_kernel<BN> (tcgen05 GEMM + implicit-GEMM conv)",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                "peak_source": peak_src, "launches_timed": int(g_n), "avg_launch_us": 1e3 * g_ms / max(1, g_n),
                "algorithmic_flops_per_launch": g_fl / max(1, g_n),
                "attention_fwd": {"achieved_tflops": (fl[2] / (ms[2] / 1e3) / 1e12) if ms[2] > 0 else None,
                                  "launches": int(cnt[2])},
                "step_level": {"algorithmic_tflop_per_image": flops_per_image(20) / 1e12,
                               "achieved_tflops_per_gpu": flops_per_image(20) * value / world / 1e12,
                               "frac_of_peak": flops_per_image(20) * value / world / 1e12 / peak_tf}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_sample()
        t_f = min(cpu_sample() for _ in range(3))
        t_fb = min(cpu_sample(backward=True) for _ in range(2))
        cpu = {"value": cpu_images_per_sec(t_f, t_fb), "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": "fp32 oracle (PyTorch restatement of the reference's diffusers path), SDXL UNet mid block "
                         "(797 GFLOP) at batch 1: forward x3, forward+backward(LoRA) x2; images/s extrapolated by FLOPs "
                         "with the reference step structure at E[n]=20 (616 TFLOP/image)",
               "t_fwd_s": t_f, "t_fwd_bwd_s": t_fb, "cpu_tflops_fwd": F_MID / t_f / 1e12,
               "host_cores": os.cpu_count()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_dict(world), "roofline": roof, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                        "ms_per_step": ms_e2e / args.steps},
                "lean": {"value": imgs / (ms_lean / 1e3), "unit": UNIT, "ms_per_step": ms_lean / args.steps,
                         "note": "same step with the generator objective elided on the discriminator turn, where the "
                                 "reference recomputes and discards it (output-preserving: identical loss_D / updates, "
                                 "tests/test_flash_step_cpu.py::test_lean_discriminator_turn_is_output_preserving); "
                                 "algorithmic FLOPs per image drop from 616 to ~330 TFLOP"},
                "gpu_launches": int(launches), "clocks": clocks,
                "losses_last_step": sink[-1] if sink else None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
