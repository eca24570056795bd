#!/usr/bin/env python
"""bench.py — Flash-Diffusion distillation-step throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--config sdxl|sd15|pixart|sd3|sample]     our arm (N>1: torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W                            reference arm (host CPU)

Default workload = BASELINE config 2 (config.workload): SDXL UNet 1024x1024 (latent 128x128) LoRA-rank-64 distillation
step, batch 4 per GPU, bf16, K=32 trailing DPM-Solver++ teacher, DMD + lsgan, l2 distill, synthetic latents / text
embeddings, random-init weights (seed 1234).  A "step" is one full `TrainingPipeline.training_step` (both optimizer
turns, reference src/flash/trainer/trainer.py:169-218).  The teacher-rollout length depends on the sampled start index
(flash_diffusion_model.py:167,289); the timed steps pin start_idx to the four mixture modes in turn (order 0, 3K/4, K/4, K/2)
(uniform weights = stage 3 of flash_sdxl.yaml:27-32, E[n] = 5K/8) so every run does the same work.
`--config sd15|pixart|sd3` runs BASELINE configs 1 / 3 / 4 the same way (batch = the example yaml's BATCH_SIZE);
`--config sample` is config 5: latency of `sample(num_steps=4)` at batch 1..32 plus achieved HBM GB/s.

The reference (diffusers fork + peft + lightning) is not installable offline, so the reference arm times
the fp32 ORACLE — the restatement of the reference's diffusers path — on the host cores: per step, one teacher forward,
one LoRA student forward+backward and one GAN-backbone forward+backward of the SDXL oracle at batch 1, combined with the
reference step's call counts into images/s (SURVEY.md §8d "CPU baseline"), plus the FULL config-1 step run for real.

One JSON line on stdout (rank 0).  See DESIGN.md §6 for every key.
"""
import argparse
import ctypes
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
F_FWD, F_DM = 6.76e12, 2.93e12              # SDXL UNet FLOPs / sample: full forward, down+mid only
# LoRA weight gradients per sample: sum over the 560 adapted linears of 2*tok*r*(K+N) (dA = dt^T x, dB = dy^T t):
# 360 linears at 1024 tok x 1280 ch, 60 at 4096 tok x 640 ch, 140 cross-attention k/v at 77 tok -> 0.17 TFLOP (0.03 %)
F_LORA_DW = 0.17e12


def flops_per_image(n):
    """2*[(4+2n)F + 2 F_dm] + F (student dX) + F_dm (GAN dX through the teacher) + 2 F_lora_dW."""
    return 2 * ((4 + 2 * n) * F_FWD + 2 * F_DM) + F_FWD + F_DM + 2 * F_LORA_DW


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


# ------------------------------------------------------------------------------------------------ workloads
def _cfg(name):
    from flash import recipes
    if name == "sdxl":
        return dict(B=4, K=32, build=recipes.build_sdxl_distillation,
                    batch=lambda B, seed: recipes.synthetic_batch(B, 128, 77, 2048, 1280, seed=seed),
                    workload="Flash-SDXL UNet 1024x1024 (latent 128x128) LoRA-rank-64 distillation step (student fwd+bwd, "
                             "K=32 DPM-Solver++ teacher CFG rollout, DMD, lsgan GAN), bf16, batch 4/GPU",
                    extra={"latent": [4, 128, 128], "context": [77, 2048], "vector": 2816, "lora_rank": 64})
    if name == "sd15":
        return dict(B=4, K=32, build=recipes.build_sd15_distillation,
                    batch=lambda B, seed: recipes.synthetic_batch(B, 64, 77, 768, 0, seed=seed, image_px=512.0),
                    workload="Flash-SD1.5 UNet 512x512 (latent 64x64) LoRA-rank-128 distillation step (K=32 DPM-Solver++ "
                             "teacher CFG rollout, DMD, lsgan), bf16, batch 4/GPU (BASELINE config 1 on the GPU kernels)",
                    extra={"latent": [4, 64, 64], "context": [77, 768], "lora_rank": 128})
    if name == "pixart":
        return dict(B=2, K=16, build=recipes.build_pixart_distillation,
                    batch=lambda B, seed: recipes.pixart_batch(B, seed, "cpu"),
                    workload="Flash-PixArt-alpha XL/2 DiT 1024x1024 (latent 128x128) LoRA-rank-64 distillation step (K=16 "
                             "DPM-Solver++ teacher CFG rollout, DMD, lsgan), bf16, batch 2/GPU (BASELINE config 3)",
                    extra={"latent": [4, 128, 128], "context": [120, 4096], "lora_rank": 64})
    if name == "sd3":
        return dict(B=2, K=32, build=recipes.build_sd3_distillation,
                    batch=lambda B, seed: recipes.sd3_batch(B, seed, "cpu"),
                    workload="Flash-SD3-medium MMDiT 1024x1024 (latent 16x128x128) LoRA-rank-64 distillation step (K=32 "
                             "flow-matching Euler teacher CFG rollout, DMD, lsgan), bf16, batch 2/GPU (BASELINE config 4; "
                             "LoRA rank = examples/configs/flash_sd3.yaml)",
                    extra={"latent": [16, 128, 128], "context": [154, 4096], "lora_rank": 64})
    raise ValueError(name)


def config_dict(name, cfg, n_gpus):
    K = cfg["K"]
    d = {"workload": cfg["workload"], "name": name, "global_batch": cfg["B"] * n_gpus, "batch_per_gpu": cfg["B"], "K": K,
         "start_idx_schedule": [0, 3 * K // 4, K // 4, K // 2], "expected_teacher_steps": 5 * K // 8,
         "parallelism": f"dp{n_gpus}", "l2_policy": "inputs and activations larger than L2 (bf16 weights alone are GBs)",
         "weights": "random-init seed 1234", "distill_loss": "l2"}
    d.update(cfg["extra"])
    return d


# ------------------------------------------------------------------------------------------------ CPU baseline
# The reference's diffusers/peft path cannot be installed: the fp32 ORACLE restates it.  A full SDXL
# step is ~2.5 PFLOP per batch (hours on CPU), so the three DISTINCT denoiser passes of the step are timed at batch 1
# on the SDXL architecture and combined with the reference step's call counts, E[n] = 20:
#   per image and training_step (two optimizer turns, each a full forward, trainer.py:169-218):
#     teacher forward (no grad)      2 x (2n CFG rollout + 2 DMD real)            = 84
#     student forward (no grad, DMD) 2 x 1                                        = 2     (timed as a teacher forward)
#     student forward+backward       forward in both turns, backward in turn 0    = 1 fwd+bwd + 1 fwd
#     GAN backbone (down+mid) at 2B  2 x 2 forwards, backward of the fake half in turn 0 = 1 fwd+bwd + 3 fwd
_CPU = {}


def cpu_threads():
    # physical cores of the host (2 hyper-threads per core on the pool's Xeon boxes); pinned so that runs agree
    return max(1, min((os.cpu_count() or 2) // 2, 64))


def _cpu_sdxl_student():
    if "net" not in _CPU:
        from oracle.unet import SDXL_KWARGS, LoraConfig, UNet2DConditionOracle
        torch.manual_seed(1234)
        with torch.device("meta"):
            net = UNet2DConditionOracle(**SDXL_KWARGS)
        net = net.to_empty(device="cpu")
        with torch.no_grad():
            for p in net.parameters():
                p.normal_(0.0, 0.02) if p.dim() >= 2 else p.fill_(1.0 if p.dim() == 1 and p.numel() > 4 else 0.0)
        for p in net.parameters():
            p.requires_grad = False
        net.add_adapter(LoraConfig(r=64, lora_alpha=64, init_lora_weights="gaussian",
                                   target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
        _CPU["net"] = net
    return _CPU["net"]


def _cpu_inputs():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 4, 128, 128, generator=g)
    t = torch.tensor([500.0])
    cond = {"cond": {"crossattn": torch.randn(1, 77, 2048, generator=g), "vector": torch.randn(1, 2816, generator=g)}}
    return x, t, cond


def cpu_teacher_forward():
    """seconds of ONE SDXL forward (no grad) at batch 1 on the host cores — the pass a reference step runs 87 times
    per image."""
    torch.set_num_threads(cpu_threads())
    net = _cpu_sdxl_student()
    x, t, cond = _cpu_inputs()
    with torch.no_grad():
        t0 = time.perf_counter()
        net(x, t, cond)
        return time.perf_counter() - t0


def cpu_mid_block():
    """seconds of ONE forward of the SDXL UNet mid block (Res + Transformer2D depth 10 at 1024 tokens x 1280 channels +
    Res, 0.8 TFLOP) at batch 1 — the short per-step sample of the reference arm: it tracks the host's drift between
    steps, the absolute numbers come from the four full passes timed up front."""
    torch.set_num_threads(cpu_threads())
    net = _cpu_sdxl_student()
    g = torch.Generator().manual_seed(11)
    x, temb, ctx = torch.randn(1, 1280, 32, 32, generator=g), torch.randn(1, 1280, generator=g), torch.randn(1, 77, 2048, generator=g)
    with torch.no_grad():
        t0 = time.perf_counter()
        net.mid_block(x, temb, ctx)
        return time.perf_counter() - t0


def cpu_components(repeats=1):
    """seconds (min over `repeats`) of the four distinct SDXL passes at batch 1 on the host cores."""
    torch.set_num_threads(cpu_threads())
    net = _cpu_sdxl_student()
    x, t, cond = _cpu_inputs()
    out = {"t_teacher_fwd": 1e30, "t_student_fwd_bwd": 1e30, "t_backbone_fwd": 1e30, "t_backbone_fwd_bwd": 1e30}
    for _ in range(repeats):
        out["t_teacher_fwd"] = min(out["t_teacher_fwd"], cpu_teacher_forward())
        with torch.no_grad():
            t0 = time.perf_counter()
            net(x, t, cond, return_intermediate=True)
            out["t_backbone_fwd"] = min(out["t_backbone_fwd"], time.perf_counter() - t0)
        t0 = time.perf_counter()
        net(x, t, cond).square().mean().backward()
        out["t_student_fwd_bwd"] = min(out["t_student_fwd_bwd"], time.perf_counter() - t0)
        xg = x.clone().requires_grad_(True)
        t0 = time.perf_counter()
        net(xg, t, cond, return_intermediate=True).square().mean().backward()
        out["t_backbone_fwd_bwd"] = min(out["t_backbone_fwd_bwd"], time.perf_counter() - t0)
        for p in net.parameters():
            p.grad = None
    return out


def cpu_images_per_sec(c, n=20):
    per_image = ((2 * (2 * n + 2) + 2 + 1) * c["t_teacher_fwd"] + c["t_student_fwd_bwd"]
                 + 3 * c["t_backbone_fwd"] + c["t_backbone_fwd_bwd"])
    return 1.0 / per_image, per_image


def cpu_config1_full_step():
    """BASELINE config 1 run FOR REAL on the host: SD1.5 architecture 512x512, random init, batch 1, fp32 oracle
    denoisers inside the product's FlashDiffusion / TrainingPipeline host logic, K = 4."""
    import copy
    from flash import recipes
    from flash.models.flash import FlashDiffusion, FlashDiffusionConfig
    from flash.schedulers import DPMSolverMultistepScheduler, LCMScheduler
    from flash.trainer import TrainingConfig, TrainingPipeline
    from oracle.unet import SD15_KWARGS, LoraConfig, UNet2DConditionOracle
    torch.set_num_threads(cpu_threads())
    torch.manual_seed(0)
    teacher = UNet2DConditionOracle(**SD15_KWARGS)
    student = copy.deepcopy(teacher)
    student.add_adapter(LoraConfig(r=128, lora_alpha=128, init_lora_weights="gaussian",
                                   target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
    with torch.no_grad():
        for n_, p_ in student.named_parameters():
            if "lora_B" in n_:
                p_.normal_(0.0, 0.02)
    teacher.freeze()
    cfg = FlashDiffusionConfig(K=[4], num_iterations_per_K=[10 ** 9], guidance_scale_min=3.0, guidance_scale_max=13.0,
                               distill_loss_type="l2", ucg_keys=["text_emb"], use_dmd_loss=True, gan_loss_type="lsgan",
                               timestep_distribution="mixture", mixture_num_components=4, mixture_var=0.5,
                               input_key="image")
    repo = "stabilityai/stable-diffusion-xl-base-1.0"
    model = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher,
                           teacher_noise_scheduler=DPMSolverMultistepScheduler.from_pretrained(repo, timestep_spacing="trailing"),
                           sampling_noise_scheduler=LCMScheduler.from_pretrained(repo, timestep_spacing="trailing"),
                           vae=None, conditioner=recipes.text_only_conditioner(),
                           discriminator=recipes.sd15_discriminator())
    pipe = TrainingPipeline(model, TrainingConfig(optimizers_name=["AdamW", "AdamW"], learning_rates=[1e-5, 1e-5],
                                                  trainable_params=[["student_denoiser"], ["discriminator."]]))
    pipe.configure_optimizers()
    batch = recipes.synthetic_batch(1, 64, 77, 768, 0, seed=1, image_px=512.0)
    t0 = time.perf_counter()
    out = pipe.training_step(batch, 0, draws={"start_idx": 0})
    dt = time.perf_counter() - t0
    return {"seconds": dt, "images_per_s": 1.0 / dt, "K": 4, "teacher_steps": 4, "batch": 1,
            "losses": [float(out["loss_optimizer_0"]), float(out["loss_optimizer_1"])]}


def cpu_block(c, kind_note):
    v, per_image = cpu_images_per_sec(c)
    return {"value": v, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "host_cpus": os.cpu_count(),
            "extrapolated": True, "seconds_per_image": per_image,
            "sample": "fp32 oracle (PyTorch restatement of the reference's diffusers/peft path) of the SDXL UNet at batch "
                      "1 on the host cores: teacher forward, LoRA student forward+backward, GAN backbone (down+mid) "
                      "forward and forward+backward, " + kind_note + "; images/s = 1 / sum(count_i * t_i) with the "
                      "reference step's call counts at E[n]=20 (87 fwd + 1 student fwd+bwd + 3 backbone fwd + 1 backbone "
                      "fwd+bwd per image)",
            **{k: round(v_, 4) for k, v_ in c.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != "sdxl":
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU reference arm covers the headline config (sdxl); "
                                                              f"--config {args.config} has no CPU arm"}))
        return
    t_all = time.perf_counter()
    cpu_teacher_forward()                              # build + first touch outside the timed samples
    once = cpu_components(1)                           # the four full SDXL passes at batch 1: timed once, up front
    once["t_teacher_fwd"] = min(once["t_teacher_fwd"], cpu_teacher_forward())
    cpu_mid_block()
    t_mid0 = min(cpu_mid_block() for _ in range(3))
    per_image0 = cpu_images_per_sec(once)[1]
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        tm = cpu_mid_block()                           # the bounded per-step sample (host drift between steps)
        if i >= args.warmup:
            walls.append(time.perf_counter() - t0)
            vals.append(1.0 / (per_image0 * tm / t_mid0))
    v = max(vals)                                      # = the least-disturbed sample (min time)
    per_image = 1.0 / v
    cfg = _cfg("sdxl")
    full1 = cpu_config1_full_step() if args.with_config1 or args.steps <= 8 else None
    sample_s = sum(walls) / len(walls)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sample_s, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict("sdxl", cfg, args.gpus),
            "sample_fraction_of_a_step": sample_s / (per_image * cfg["B"]),
            "cpu_baseline": dict(cpu_block(once, "each timed once up front (teacher forward: min of 2)"),
                                 value=v, seconds_per_image=per_image, t_mid_block_upfront=t_mid0,
                                 per_step_values=vals, config1_full_step=full1),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
            "note": "per step: ONE forward of the SDXL oracle's mid block at batch 1 (0.8 TFLOP, ms_per_step = its wall "
                    "time) rescales the four full batch-1 passes timed up front (teacher forward, student forward+backward, "
                    "GAN backbone forward and forward+backward); value = the images/s a full reference step reaches at "
                    "those per-pass times (a full step of batch 4 = 1 / sample_fraction_of_a_step samples); "
                    "config1_full_step = BASELINE config 1 run for real on the CPU (with --with-config1 or <= 8 steps)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ same-box eager bar
def torch_eager_block(dev):
    """The same-box "kernel to beat" (BASELINE.md §3): what today's torch eager (cuBLASLt / cuDNN /
    SDPA under bf16 autocast) does with the SDXL oracle modules on THIS GPU: one 2B teacher evaluation (batch 8, no
    grad) and one LoRA student forward+backward (batch 4)."""
    from oracle.unet import SDXL_KWARGS, LoraConfig, UNet2DConditionOracle

    def timeit(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    torch.manual_seed(0)
    with torch.device(dev):
        net = UNet2DConditionOracle(**SDXL_KWARGS)
    for p in net.parameters():
        p.requires_grad = False
    net.add_adapter(LoraConfig(r=64, lora_alpha=64, init_lora_weights="gaussian",
                               target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
    net = net.to(dev)
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(8, 4, 128, 128, device=dev, generator=g)
    t = torch.full((8,), 500.0, device=dev)
    cond = {"cond": {"crossattn": torch.randn(8, 77, 2048, device=dev, generator=g),
                     "vector": torch.randn(8, 2816, device=dev, generator=g)}}

    def teacher():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            net(x, t, cond)

    def student():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = net(x[:4], t[:4], {"cond": {k: v[:4] for k, v in cond["cond"].items()}})
        y.float().square().mean().backward()

    ms_t, ms_s = timeit(teacher), timeit(student)
    del net
    torch.cuda.empty_cache()
    return {"teacher_eval_batch8_ms": ms_t, "teacher_eval_tflops": 8 * F_FWD / ms_t / 1e9,
            "student_fwd_bwd_batch4_ms": ms_s,
            "what": "fp32 oracle modules (restated diffusers UNet2DConditionModel + peft LoRA) under torch.autocast(bf16), "
                    "torch eager: cuBLASLt GEMMs, cuDNN convs, F.scaled_dot_product_attention"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    from flash.b200 import graphs
    from flash.b200 import lib as fdlib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = fdlib.load()
    lib.fd_launch_count.restype = ctypes.c_longlong
    cfg = _cfg(args.config)
    B, K = cfg["B"], cfg["K"]
    # the four mixture modes in an order whose running mean of teacher steps (K - start_idx) is 5K/8 after every EVEN
    # number of steps, so a --steps count that is not a multiple of 4 still times the expected rollout length
    modes = [0, 3 * K // 4, K // 4, K // 2]
    model, pipe = cfg["build"](dev)

    def host_batch(i):
        return {k: v.pin_memory() for k, v in cfg["batch"](B, 1234 + rank + 1000 * i).items()}

    def step(batch, i):
        return pipe.training_step(batch, i, draws={"start_idx": modes[i % 4]})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: every start index once (each rollout length captures / warms its own code paths), then the rest
    for w in range(max(args.warmup, 3)):
        step({k: v.to(dev) for k, v in host_batch(-1 - w).items()}, (3 - w) % 4)
    resident = [{k: v.to(dev) for k, v in host_batch(i).items()} for i in range(args.steps)]

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
    ms_lean = None
    if hasattr(model, "elide_unused_generator_pass"):
        model.elide_unused_generator_pass = True
        n_lean = min(args.steps, 4)
        ms_lean, _ = timed(lambda: [step(resident[i], i) for i in range(n_lean)])
        model.elide_unused_generator_pass = False
    clocks = sampler.stop() if rank == 0 else None
    ar = getattr(pipe, "allreduce_stats", None)

    imgs = B * world * args.steps
    value = imgs / (ms_dev / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)

    # roofline: ONE MORE FULL training step (start_idx = 3K/4), launched eagerly instead of from CUDA graphs so that
    # every tensor-core launch of the step — teacher rollout, DMD, GAN backbone, LoRA student forward AND backward —
    # sits between two CUDA events on its launch stream (fd_profile_enable).  Same kernels, same shapes, same data as
    # the timed steps; the graph replays of the timed steps cannot carry per-kernel events.
    # EVERY rank runs this step (its gradient all-reduce is a collective); only rank 0 carries the event pairs.
    roof = None
    had = getattr(model, "use_cuda_graphs", None)
    if had is not None:
        model.use_cuda_graphs = False
    if rank == 0:
        lib.fd_profile_enable(1)
    step(resident[0], modes.index(3 * K // 4))       # the start_idx = 3K/4 mode, whatever its place in the timed order
    torch.cuda.synchronize()
    if rank == 0:
        lib.fd_profile_enable(0)
    if had is not None:
        model.use_cuda_graphs = had
    if rank == 0:
        peak_tf, peak_hbm, peak_src = peaks()
        dump = os.path.join(ROOT, "gpurun_out", f"bench_launches_{args.config}.csv")
        try:
            os.makedirs(os.path.dirname(dump), exist_ok=True)
            lib.fd_profile_dump(dump.encode())
        except Exception:
            dump = None
        ms = (ctypes.c_double * 4)()
        fl = (ctypes.c_double * 4)()
        cnt = (ctypes.c_longlong * 4)()
        lib.fd_profile_summary(ms, fl, cnt, 4)
        g_ms, g_fl, g_n = ms[0] + ms[1], fl[0] + fl[1], cnt[0] + cnt[1]
        # the r01 definition, for comparison: the GEMM / conv launches of ONE 2B teacher evaluation (the op that is
        # ~83 % of a step), eager, same event pairs
        teacher_eval = None
        if hasattr(model, "_teacher_pair") and hasattr(model, "conditioner") and args.config == "sdxl":
            if had is not None:
                model.use_cuda_graphs = False
            lib.fd_profile_enable(1)
            with torch.no_grad():
                b0 = resident[0]
                cond = model.conditioner(b0, set_ucg_rate_zero=True)
                unc = model.conditioner(b0, ucg_keys=model.ucg_keys)
                model.__dict__["_kv_key"] = None
                model._teacher_pair(model.teacher_denoiser, b0["image"], torch.full((B,), 500, device=dev), cond, unc)
            torch.cuda.synchronize()
            lib.fd_profile_enable(0)
            if had is not None:
                model.use_cuda_graphs = had
            ms2 = (ctypes.c_double * 4)()
            fl2 = (ctypes.c_double * 4)()
            cnt2 = (ctypes.c_longlong * 4)()
            lib.fd_profile_summary(ms2, fl2, cnt2, 4)
            t_ms, t_fl = ms2[0] + ms2[1], fl2[0] + fl2[1]
            if t_ms > 0:
                teacher_eval = {"achieved_tflops": t_fl / (t_ms / 1e3) / 1e12, "frac": t_fl / (t_ms / 1e3) / 1e12 / peak_tf,
                                "launches": int(cnt2[0] + cnt2[1]),
                                "attention_fwd_tflops": (fl2[2] / (ms2[2] / 1e3) / 1e12) if ms2[2] > 0 else None}
        ach = g_fl / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                traffic = json.load(f)
        except Exception:
            pass
        n_exp = 5 * K // 8
        step_flops_profiled = fl[0] + fl[1] + fl[2] + fl[3]
        roof = {"bound": "tensor", "kernel": "fd::gemm_pair_kernel<BN> (tcgen05 cta_group::2 stream-K GEMM + implicit-GEMM conv)",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": (traffic or {}).get("dominant_dram_bytes_per_launch"), "traffic_detail": traffic,
                "peak_source": peak_src, "launches_timed": int(g_n), "avg_launch_us": 1e3 * g_ms / max(1, g_n),
                "algorithmic_flops_per_launch": g_fl / max(1, g_n),
                "measured_in": "one extra full training step (start_idx 3K/4) launched eagerly with a CUDA-event pair "
                               "around every tensor-core launch",
                "conv": {"achieved_tflops": (fl[1] / (ms[1] / 1e3) / 1e12) if ms[1] > 0 else None, "launches": int(cnt[1])},
                "attention_fwd": {"achieved_tflops": (fl[2] / (ms[2] / 1e3) / 1e12) if ms[2] > 0 else None,
                                  "launches": int(cnt[2]), "frac": (fl[2] / (ms[2] / 1e3) / 1e12 / peak_tf) if ms[2] > 0 else None},
                "attention_bwd": {"achieved_tflops": (fl[3] / (ms[3] / 1e3) / 1e12) if ms[3] > 0 else None,
                                  "launches": int(cnt[3])},
                "teacher_evaluation": teacher_eval,
                "tensor_core_flops_in_profiled_step": step_flops_profiled,
                "per_launch_csv": os.path.relpath(dump, ROOT) if dump else None}
        if args.config == "sdxl":
            roof["step_level"] = {"algorithmic_tflop_per_image": flops_per_image(n_exp) / 1e12,
                                  "achieved_tflops_per_gpu": flops_per_image(n_exp) * value / world / 1e12,
                                  "frac_of_peak": flops_per_image(n_exp) * value / world / 1e12 / peak_tf}

    eager = None
    if rank == 0 and world == 1 and args.config == "sdxl" and not args.no_eager_baseline:
        del resident, hosts
        torch.cuda.empty_cache()
        try:
            eager = torch_eager_block(dev)
        except Exception as e:          # the bar is informative; never lose the headline line over it
            eager = {"error": repr(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and args.config == "sdxl" and not args.no_cpu_baseline:
        # bounded (~25 s of CPU work after the build): the teacher forward and the GAN backbone forward are timed; the
        # two forward+backward passes are taken as 3.0x / 2.0x their forwards (ratios the reference arm measures)
        cpu_teacher_forward()
        t_f = cpu_teacher_forward()
        net = _cpu_sdxl_student()
        x_, t_, c_ = _cpu_inputs()
        with torch.no_grad():
            t0 = time.perf_counter()
            net(x_, t_, c_, return_intermediate=True)
            t_b = time.perf_counter() - t0
        cpu = cpu_block({"t_teacher_fwd": t_f, "t_student_fwd_bwd": 3.0 * t_f, "t_backbone_fwd": t_b,
                         "t_backbone_fwd_bwd": 2.0 * t_b},
                        "the two forwards timed once; forward+backward passes taken as 3.0x / 2.0x their forwards "
                        "(`bench.py --impl reference` times all four)")

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_dict(args.config, cfg, world), "roofline": roof, "cpu_baseline": cpu,
                "torch_eager_b200": eager,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": clocks,
                "allreduce": ar,
                "losses_last_step": sink[-1] if sink else None,
                "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        if ms_lean is not None:
            line["lean"] = {"value": B * world * n_lean / (ms_lean / 1e3), "unit": UNIT, "ms_per_step": ms_lean / n_lean,
                            "steps": n_lean,
                            "note": "same step with the generator objective elided on the discriminator turn, where the "
                                    "reference recomputes and discards it (output-preserving: identical loss_D / updates, "
                                    "tests/test_flash_step_cpu.py::test_lean_discriminator_turn_is_output_preserving)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ config 5
def run_sample(args):
    """BASELINE config 5: 4-NFE `sample()` latency (reference flash_diffusion_model.py:754-915; cond + uncond evaluated
    every step as the reference does, no VAE decode) at batch 1..32 on one GPU, with the achieved HBM GB/s: algorithmic
    bytes = the student's bf16 weights streamed once per denoiser call (4 calls at batch 2B) + the latents in and out."""
    from flash import recipes
    from flash.b200 import graphs
    from flash.b200 import lib as fdlib
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lib = fdlib.load()
    lib.fd_launch_count.restype = ctypes.c_longlong
    peak_tf, peak_hbm, peak_src = peaks()
    which = args.backbone
    if which == "sdxl":
        model, _ = recipes.build_sdxl_distillation(dev)
        mk = lambda B: recipes.synthetic_batch(B, 128, 77, 2048, 1280, seed=B, device=dev)
        shape, fwd = (4, 128, 128), 6.76e12
    elif which == "sd15":
        model, _ = recipes.build_sd15_distillation(dev)
        mk = lambda B: recipes.synthetic_batch(B, 64, 77, 768, 0, seed=B, device=dev, image_px=512.0)
        shape, fwd = (4, 64, 64), 0.80e12
    elif which == "pixart":
        model = recipes.build_pixart_sampler(dev)
        mk = lambda B: recipes.pixart_batch(B, B, dev)
        shape, fwd = (4, 128, 128), 6.51e12
    else:
        model, _ = recipes.build_sd3_distillation(dev)
        mk = lambda B: recipes.sd3_batch(B, B, dev)
        shape, fwd = (16, 128, 128), 8.4e12
    model.eval()
    wbytes = 2 * sum(p.numel() for n, p in model.student_denoiser.named_parameters())     # bf16 packs of every weight
    rows = []
    l0 = lib.fd_launch_count() + graphs.REPLAYED_LAUNCHES
    for B in [1, 2, 4, 8, 16, 32]:
        batch = mk(B)
        z = torch.randn(B, *shape, device=dev)
        try:
            for _ in range(max(args.warmup, 3)):
                model.sample(z, num_steps=4, guidance_scale=1.0, conditioner_inputs=batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(args.steps, 3)
            e0.record()
            for _ in range(n):
                model.sample(z, num_steps=4, guidance_scale=1.0, conditioner_inputs=batch)
            e1.record()
            torch.cuda.synchronize()
        except torch.cuda.OutOfMemoryError:
            break
        ms = e0.elapsed_time(e1) / n
        calls = 4                                   # one (2B- or B-batched) denoiser call per step
        bytes_alg = calls * wbytes + calls * 2 * (2 * B) * shape[0] * shape[1] * shape[2] * 4
        # denoiser evaluations per sample(): 4 steps x (cond + uncond) as the reference's loop does; the SD3 class skips the
        # unconditional branch when guidance_scale == 1 (it is multiplied by exactly 0): 4 evaluations
        n_eval = 4 if which == "sd3" else 8
        rows.append({"batch": B, "latency_ms": ms, "images_per_s": B / ms * 1e3, "tflops": n_eval * B * fwd / ms / 1e9,
                     "denoiser_evaluations": n_eval,
                     "hbm_gbs": bytes_alg / ms / 1e6, "hbm_frac_of_peak": bytes_alg / ms / 1e6 / peak_hbm})
    launches = lib.fd_launch_count() + graphs.REPLAYED_LAUNCHES - l0
    r1 = rows[0]
    print(json.dumps({"metric": "4-NFE sample latency", "value": r1["latency_ms"], "unit": "ms", "n_gpus": 1,
                      "steps": max(args.steps, 3), "warmup": max(args.warmup, 3), "ms_per_step": r1["latency_ms"],
                      "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                      "data": "synthetic",
                      "config": {"workload": f"4-NFE sampler (LCM / flow-matching student loop, cond+uncond per step, no "
                                             f"VAE decode), {which} LoRA student, batch 1 (rows: batch 1..32)",
                                 "backbone": which, "latent": list(shape), "weights": "random-init seed 1234"},
                      "roofline": {"bound": "tensor", "achieved": r1["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                                   "frac": r1["tflops"] / peak_tf, "traffic": None, "peak_source": peak_src,
                                   "hbm_gbs_batch1": r1["hbm_gbs"], "hbm_peak_gbs": peak_hbm,
                                   "note": "even at batch 1 the sampler is tensor-bound: streaming the bf16 weights once "
                                           "per call needs hbm_gbs_batch1 of the measured HBM peak"},
                      "rows": rows, "gpu_launches": int(launches), "cpu_baseline": None,
                      "e2e": {"value": r1["latency_ms"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                              "note": "latents are generated on the device, as the reference's log_samples does"}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="sdxl", choices=["sdxl", "sd15", "pixart", "sd3", "sample"])
    ap.add_argument("--backbone", default="sdxl", choices=["sdxl", "sd15", "pixart", "sd3"], help="for --config sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--with-config1", action="store_true",
                    help="reference arm: also run the full config-1 step on the CPU (~90 s; default for <= 8 steps)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "sample":
        run_sample(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
