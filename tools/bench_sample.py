"""4-NFE sampler latency (BASELINE config 5, SDXL row): python tools/bench_sample.py
`FlashDiffusion.sample(num_steps=4)` (reference flash_diffusion_model.py:754-915) without VAE decode, batch 1..32."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash import recipes

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
if which == "sdxl":
    model, _ = recipes.build_sdxl_distillation(dev)
    hw, ctx, pooled, fwd_flops = 128, 2048, 1280, 6.76e12
elif which == "pixart":
    model = recipes.build_pixart_sampler(dev)
    hw, ctx, pooled, fwd_flops = 128, 4096, 0, 6.51e12
else:
    model, _ = recipes.build_sd15_distillation(dev)
    hw, ctx, pooled, fwd_flops = 64, 768, 0, 0.80e12
model.eval()
rows = []
for B in [1, 2, 4, 8, 16, 32]:
    batch = (recipes.pixart_batch(B, B, dev) if which == "pixart"
             else recipes.synthetic_batch(B, hw, 77, ctx, pooled, seed=B, device=dev))
    z = torch.randn(B, 4, hw, hw, device=dev)
    for _ in range(2):
        model.sample(z, num_steps=4, guidance_scale=1.0, conditioner_inputs=batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        out, _ = model.sample(z, num_steps=4, guidance_scale=1.0, conditioner_inputs=batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 8 * B * fwd_flops        # 4 steps x (cond + uncond) student evaluations, as the reference does
    rows.append({"batch": B, "latency_ms": ms, "images_per_s": B / ms * 1e3, "tflops": flops / ms / 1e9})
    print(rows[-1], flush=True)
print(json.dumps({"metric": f"4-NFE sample latency ({which} student, LoRA, CFG cond+uncond as reference)", "rows": rows}))
