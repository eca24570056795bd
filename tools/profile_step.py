"""Kernel-time breakdown of one full SDXL distillation step with torch.profiler (CUPTI, no replay):
python tools/profile_step.py [start_idx] > gpurun_out/step_profile.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch
from torch.profiler import ProfilerActivity, profile

from flash import recipes

start_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
model, pipe = recipes.build_sdxl_distillation(dev)
batch = recipes.synthetic_batch(4, 128, 77, 2048, 1280, seed=1, device=dev)
for _ in range(2):
    pipe.training_step(batch, 0, draws={"start_idx": 24})
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pipe.training_step(batch, 0, draws={"start_idx": start_idx})
e1.record(); torch.cuda.synchronize()
print(f"step (start_idx={start_idx}, n={32 - start_idx}) : {e0.elapsed_time(e1):.1f} ms without profiler")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pipe.training_step(batch, 0, draws={"start_idx": start_idx})
    torch.cuda.synchronize()
rows = {}
total = 0.0
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0)
    if t > 0 and ev.device_type.name == "CUDA" if hasattr(ev, "device_type") else t > 0:
        rows[ev.key] = (t, ev.count)
        total += t
print(f"total device kernel time {total / 1e3:.1f} ms")
for k, (t, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{100 * t / total:6.2f}%  {t / 1e3:9.2f} ms  {c:7d}  {k[:100]}")
