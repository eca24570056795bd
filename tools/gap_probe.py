"""Where the GPU idles inside one SDXL distillation step: kernel timeline from torch.profiler (CUPTI), gaps between
consecutive kernels binned by size and by what runs next.  python tools/gap_probe.py [start_idx]"""
import collections
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
for _ in range(3):
    pipe.training_step(batch, 0, draws={"start_idx": 24})
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe.training_step(batch, 0, draws={"start_idx": start_idx})
    torch.cuda.synchronize()
evs = []
for e in prof.events():
    if e.device_type.name == "CUDA" and e.time_range is not None:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
t0, t1 = evs[0][0], max(e[1] for e in evs)
busy_end = evs[0][1]
busy = evs[0][1] - evs[0][0]
bins = [(2, "<2us"), (5, "2-5us"), (20, "5-20us"), (100, "20-100us"), (1e9, ">100us")]
hist = collections.OrderedDict((b[1], [0, 0.0]) for b in bins)
after = collections.defaultdict(lambda: [0, 0.0])
gaps = []
for (s, e, name) in evs[1:]:
    if s > busy_end:
        g = s - busy_end
        for lim, label in bins:
            if g < lim:
                hist[label][0] += 1
                hist[label][1] += g
                break
        key = name.split("(")[0][:60]
        after[key][0] += 1
        after[key][1] += g
        gaps.append((g, busy_end - t0, name[:70]))
        busy += e - s
        busy_end = e
    elif e > busy_end:
        busy += e - busy_end
        busy_end = e
span = t1 - t0
print(f"kernels {len(evs)}  span {span / 1e3:.1f} ms  busy {busy / 1e3:.1f} ms  idle {(span - busy) / 1e3:.1f} ms ({100 * (span - busy) / span:.1f}%)")
for k, (n, t) in hist.items():
    print(f"  gaps {k:>9}: {n:7d}  {t / 1e3:8.2f} ms")
print("idle time by the kernel that ends the gap:")
for k, (n, t) in sorted(after.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {t / 1e3:8.2f} ms  {n:7d} gaps  avg {t / n:6.2f} us  {k}")
print("largest gaps (ms into the step):")
for g, at, name in sorted(gaps, reverse=True)[:12]:
    print(f"  {g:9.1f} us at {at / 1e3:8.1f} ms before {name}")
