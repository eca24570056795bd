"""Timing of one PixArt-alpha XL/2 DiT evaluation at 1024x1024 on the B200 engine: python tools/bench_dit.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.models.transformers import DiffusersTransformer2DWrapper
from oracle.dit import PIXART_KWARGS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2 and sys.argv[2] == "sd3":
    from flash.models.transformers import DiffusersSD3Transformer2DWrapper
    from oracle.sd3 import SD3_KWARGS
    with torch.device("meta"):
        m = DiffusersSD3Transformer2DWrapper(**SD3_KWARGS)
    m = m.to_empty(device="cuda")
    torch.manual_seed(0)
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.normal_(0, 0.02)
        m.pos_embed.pos_embed.normal_(0, 0.02)
    m.freeze()
    x = torch.randn(B, 16, 128, 128, device="cuda")
    t = torch.full((B,), 500.0, device="cuda")
    cond = {"cond": {"crossattn": torch.randn(B, 154, 4096, device="cuda"), "vector": torch.randn(B, 2048, device="cuda")}}
    rows = []
    with torch.no_grad():
        for _ in range(2):
            m(x, t, cond)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m(x, t, cond)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"SD3-medium MMDiT fwd B={B} (154 text tokens): {ms:.2f} ms  -> {8.4 * B / ms * 1e3:.0f} TFLOP/s")
        # 4-NFE flow-matching sampler loop (2 evaluations per step as the reference's sample(); Euler update)
        for Bs in [1, 2, 4, 8]:
            xs = torch.randn(Bs, 16, 128, 128, device="cuda")
            cs = {"cond": {"crossattn": torch.randn(2 * Bs, 154, 4096, device="cuda"), "vector": torch.randn(2 * Bs, 2048, device="cuda")}}
            sig = [1.0, 0.75, 0.5, 0.25, 0.0]
            def run():
                z = xs.clone()
                for i in range(4):
                    v = m(torch.cat([z, z]), torch.full((2 * Bs,), sig[i] * 1000, device="cuda"), cs)
                    v = 1.0 * v[:Bs] + 0.0 * v[Bs:]
                    z = z + (sig[i + 1] - sig[i]) * v
                return z
            run(); torch.cuda.synchronize()
            e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
            print({"batch": Bs, "latency_ms": e0.elapsed_time(e1) / 2})
    sys.exit(0)
with torch.device("meta"):
    m = DiffusersTransformer2DWrapper(**PIXART_KWARGS)
m = m.to_empty(device="cuda")
torch.manual_seed(0)
with torch.no_grad():
    for n, p in m.named_parameters():
        p.normal_(0, 0.02)
import numpy as np
from flash.models.transformers.transformers import sincos_2d
m.pos_embed.pos_embed = torch.from_numpy(sincos_2d(1152, 64, 64, 2)).float()[None].cuda()
m.freeze()
x = torch.randn(B, 4, 128, 128, device="cuda")
t = torch.full((B,), 500.0, device="cuda")
cond = {"cond": {"crossattn": torch.randn(B, 120, 4096, device="cuda"), "vector": torch.randn(B, 768, device="cuda"),
                 "attention_mask": (torch.arange(120, device="cuda")[None] < 77).long().repeat(B, 1)}}
with torch.no_grad():
    for _ in range(2):
        m(x, t, cond)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        m(x, t, cond)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
print(f"PixArt-alpha XL/2 DiT fwd B={B}: {ms:.2f} ms  -> {6.51 * B / ms * 1e3:.0f} TFLOP/s "
      f"({6.51 * B / ms / 1374.2 * 1e5:.1f}% of sustained bf16 peak)")
