"""The dominant GEMMs of an SDXL teacher evaluation in the form they have IN SITU (fused epilogues), timed back to back
so that the clocks settle at their sustained (power-capped) level:  python tools/bench_gemm_insitu.py [reps]
  ff1   8192 x 10240 x 1280  LayerNorm fold + bias + GEGLU            (16.5 % of a step)
  ff1b  32768 x 5120 x 640   same at the 640-channel level            ( 4.8 %)
  o     8192 x 1280 x 1280   bias + residual + row statistics         ( 9.6 %)
  ff2   8192 x 1280 x 5120   bias + residual + row statistics         ( 7.1 %)
  qkv   8192 x 3840 x 1280   LayerNorm fold                           ( 5.0 %)
  o640  32768 x 640 x 640    bias + residual + row statistics         ( 2.8 %)
Also the ncu target for profiles/r02_ncu_gemm_*.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.b200 import raw

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
torch.manual_seed(0)


def mk(M, N, K):
    return (torch.randn(M, K, device="cuda").bfloat16(), (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16())


def case(name):
    if name in ("ff1", "ff1b"):
        M, N, K = (8192, 10240, 1280) if name == "ff1" else (32768, 5120, 640)
        a, b = mk(M, N, K)
        bias, colsum = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
        st = torch.stack([a.float().sum(1), (a.float() ** 2).sum(1)], dim=1).contiguous()
        return (M, N, K), lambda: raw.gemm(a, b, bias=bias, geglu=True, ln=(st, colsum, K, 1e-5))
    if name == "qkv":
        M, N, K = 8192, 3840, 1280
        a, b = mk(M, N, K)
        colsum, bias = torch.randn(N, device="cuda"), torch.zeros(N, device="cuda")
        st = torch.stack([a.float().sum(1), (a.float() ** 2).sum(1)], dim=1).contiguous()
        return (M, N, K), lambda: raw.gemm(a, b, bias=bias, ln=(st, colsum, K, 1e-5))
    M, N, K = {"o": (8192, 1280, 1280), "ff2": (8192, 1280, 5120), "o640": (32768, 640, 640)}[name]
    a, b = mk(M, N, K)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    stats = torch.empty(M, 2, device="cuda")
    return (M, N, K), lambda: raw.gemm(a, b, bias=bias, residual=res, rowstats=stats)


names = only or ["ff1", "ff1b", "o", "ff2", "qkv", "o640"]
cases = {n: case(n) for n in names}
# warm every case, then a long mixed warm-up so the clocks reach their sustained level
for _ in range(3):
    for n in names:
        cases[n][1]()
torch.cuda.synchronize()
for n in names:
    (M, N, K), fn = cases[n]
    for _ in range(reps):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{n:5s} {M}x{N}x{K}: {ms * 1e3:7.1f} us  {2 * M * N * K / ms / 1e9:6.0f} TF/s")
