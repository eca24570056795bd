"""Distillation-step throughput of the non-headline training configs (BASELINE configs 1, 3, 4) on ONE GPU:
    python tools/bench_step.py {sd15|pixart|sd3} [batch] [steps] [warmup]
Same timing rules as bench.py (warm-up >= 3, CUDA events on the launch stream, inputs resident in HBM and larger than
L2 together with the activations, start index cycled over the four mixture modes so that the teacher rollout length
averages to its expectation).  These are side measurements: bench.py's headline stays the SDXL step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash import recipes
from flash.b200 import graphs
from flash.b200 import lib as fdlib

which = sys.argv[1] if len(sys.argv) > 1 else "sd3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (4 if which == "sd15" else 2)      # BATCH_SIZE of the example yaml
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
warmup = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib = fdlib.load()
lib.fd_launch_count.restype = __import__("ctypes").c_longlong

if which == "sd15":
    model, pipe = recipes.build_sd15_distillation(dev)
    K, label = 32, "Flash-SD1.5 UNet 512x512 (latent 64x64), LoRA r=128, K=32 DPM-Solver++ CFG rollout, DMD, lsgan"
    make = lambda i: recipes.synthetic_batch(B, 64, 77, 768, 0, seed=100 + i, device=dev, image_px=512.0)
elif which == "pixart":
    model, pipe = recipes.build_pixart_distillation(dev)
    K, label = 16, "Flash-PixArt-alpha XL/2 DiT 1024x1024 (latent 128x128), LoRA r=64, K=16 DPM-Solver++ CFG rollout, DMD, lsgan"
    make = lambda i: recipes.pixart_batch(B, 100 + i, dev)
else:
    model, pipe = recipes.build_sd3_distillation(dev)
    K, label = 32, "Flash-SD3-medium MMDiT 1024x1024 (latent 16x128x128), LoRA r=64, K=32 flow-matching Euler CFG rollout, DMD, lsgan"
    make = lambda i: recipes.sd3_batch(B, 100 + i, dev)
modes = [0, K // 4, K // 2, 3 * K // 4]


def step(batch, i):
    return pipe.training_step(batch, i, draws={"start_idx": modes[i % 4]})


for w in range(warmup):
    step(make(-1 - w), 3)
batches = [make(i) for i in range(steps)]
torch.cuda.synchronize()
l0 = lib.fd_launch_count() + graphs.REPLAYED_LAUNCHES
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
outs = [step(batches[i], i) for i in range(steps)]
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
launches = lib.fd_launch_count() + graphs.REPLAYED_LAUNCHES - l0
print(json.dumps({"metric": "distillation images/sec", "config": label, "value": B * steps / (ms / 1e3),
                  "unit": "images/s", "n_gpus": 1, "batch": B, "steps": steps, "warmup": warmup,
                  "ms_per_step": ms / steps, "dtype": "bf16", "data": "synthetic, random-init weights",
                  "start_idx_schedule": modes, "gpu_launches": int(launches),
                  "losses_last_step": [float(outs[-1]["loss_optimizer_0"]), float(outs[-1]["loss_optimizer_1"])],
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
