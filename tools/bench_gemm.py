"""Micro-benchmark of fd_gemm tile configurations: python tools/bench_gemm.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.b200 import raw

SHAPES = [(131072, 320, 2880), (32768, 640, 640), (32768, 1920, 640), (8192, 1280, 1280), (8192, 3840, 1280), (8192, 10240, 1280),
          (8192, 1280, 5120), (32768, 5120, 640), (32768, 640, 2560), (16384, 2560, 1280)]
CFGS = [0, 256, 512 + 128, 512 + 160, 512 + 256]
only = [int(a) for a in sys.argv[1:]]
if only:
    CFGS = only
print("config      " + "  ".join(f"{str(s):>20s}" for s in SHAPES))
for cfg in CFGS:
    row = []
    for (M, N, K) in SHAPES:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            raw.gemm(a, b, out=out, force_bn=cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            raw.gemm(a, b, out=out, force_bn=cfg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        row.append(f"{2 * M * N * K / ms / 1e9:8.0f} TF/s {ms * 1e3:6.0f}us")
    print(f"{cfg:6d}      " + "  ".join(f"{r:>20s}" for r in row))
