"""Micro-benchmark of the attention / LayerNorm kernels: python tools/bench_attn.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.b200 import raw


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (B, H, Nq, Nkv) in [(8, 20, 1024, 1024), (8, 10, 4096, 4096), (4, 20, 1024, 1024), (4, 10, 4096, 4096),
                        (8, 20, 1024, 77), (8, 10, 4096, 77), (2, 16, 16384, 16384)]:
    q = torch.randn(B, Nq, H * 64, device="cuda").bfloat16()
    k = torch.randn(B, Nkv, H * 64, device="cuda").bfloat16()
    v = torch.randn(B, Nkv, H * 64, device="cuda").bfloat16()
    ms = timeit(lambda: raw.attention_fwd(q, k, v, H))
    fl = 4.0 * B * H * Nq * Nkv * 64
    o, lse = raw.attention_fwd(q, k, v, H, need_lse=True)
    do = torch.randn_like(o)
    msb = timeit(lambda: raw.attention_bwd(q, k, v, o, lse, do, H), n=5)
    print(f"attn B={B} H={H} Nq={Nq} Nkv={Nkv}: fwd {ms * 1e3:7.1f} us {fl / ms / 1e9:7.0f} TF/s | bwd {msb * 1e3:8.1f} us "
          f"{2.5 * fl / msb / 1e9:7.0f} TF/s")
for rows, C in [(8192, 1280), (32768, 640), (4096, 1280), (16384, 640)]:
    x = torch.randn(rows, C, device="cuda").bfloat16()
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    ms = timeit(lambda: raw.layernorm_fwd(x, g, b, 1e-5))
    print(f"layernorm {rows}x{C}: {ms * 1e3:6.1f} us  {4 * rows * C / ms / 1e9:6.2f} TB/s")
for NB, HW, C in [(8, 16384, 320), (8, 4096, 640), (8, 1024, 1280), (8, 1024, 2560)]:
    x = torch.randn(NB * HW, C, device="cuda").bfloat16()
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    ms1 = timeit(lambda: raw.groupnorm_stats(x, NB, HW, C, 32, 1e-5))
    st = raw.groupnorm_stats(x, NB, HW, C, 32, 1e-5)
    ms2 = timeit(lambda: raw.groupnorm_apply(x, st, g, b, NB, HW, C, 32, True))
    print(f"groupnorm {NB}x{HW}x{C}: stats {ms1 * 1e3:6.1f} us {2 * NB * HW * C / ms1 / 1e9:5.2f} TB/s | apply {ms2 * 1e3:6.1f} us "
          f"{4 * NB * HW * C / ms2 / 1e9:5.2f} TB/s")
