"""Achieved bandwidth / throughput of the DiT backward kernels and the generic attention backward (CUDA events, L2
flushed between iterations by a 256 MB write):  python tools/bench_dit_kernels.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.b200 import raw

dev = "cuda"
flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


rows = []
B, N, C = 2, 4096, 1536                     # SD3-medium image stream at batch 2
x = torch.randn(B * N, C, device=dev).bfloat16()
dy = torch.randn(B * N, C, device=dev).bfloat16()
mod = 0.1 * torch.randn(B, 3, C, device=dev)
E = B * N * C
for name, fn, bytes_ in [
    ("fd_layernorm_modulate", lambda: raw.layernorm_modulate(x, mod[:, 0], mod[:, 1], N, 1e-6), 4 * E),
    ("fd_layernorm_modulate_bwd", lambda: raw.layernorm_modulate_bwd(x, dy, mod[:, 0], N, 1e-6), 6 * E),
    ("fd_gate_residual", lambda: raw.gate_residual(x, mod[:, 2], dy, N), 6 * E),
    ("fd_gate_bwd", lambda: raw.gate_bwd(dy, x, mod[:, 2], N), 6 * E),
    ("fd_gelu_tanh_bwd", lambda: raw.gelu_tanh_bwd(x, dy), 6 * E),
]:
    ms = timed(fn)
    rows.append({"kernel": name, "shape": [B * N, C], "ms": ms, "algorithmic_GBps": bytes_ / ms / 1e6})
    print(rows[-1], flush=True)

for (B, H, d, Nq, Nkv, label) in [(2, 16, 80, 4096, 4096, "PixArt self-attention (72 -> 80)"),
                                  (2, 16, 80, 4096, 120, "PixArt cross-attention, 120 keys masked to 77"),
                                  (4, 8, 48, 4096, 4096, "SD1.5 level 1 self-attention (40 -> 48)"),
                                  (4, 8, 80, 1024, 1024, "SD1.5 level 2 self-attention"),
                                  (4, 8, 160, 256, 256, "SD1.5 level 3 self-attention (CUDA-core passes)"),
                                  (2, 24, 64, 4250, 4250, "SD3 joint attention (tuned d=64 kernels)")]:
    q = torch.randn(B, Nq, H * d, device=dev).bfloat16()
    k = torch.randn(B, Nkv, H * d, device=dev).bfloat16()
    v = torch.randn(B, Nkv, H * d, device=dev).bfloat16()
    do = torch.randn(B, Nq, H * d, device=dev).bfloat16()
    kv_len = torch.full((B,), 77, device=dev, dtype=torch.int32) if Nkv == 120 else None
    o, lse = raw.attention_fwd(q, k, v, H, need_lse=True, head_dim=d, kv_len=kv_len)
    ms_f = timed(lambda: raw.attention_fwd(q, k, v, H, need_lse=True, head_dim=d, kv_len=kv_len))
    ms_b = timed(lambda: raw.attention_bwd(q, k, v, o, lse, do, H, head_dim=d, kv_len=kv_len))
    fl = 4.0 * B * H * Nq * Nkv * d
    rows.append({"kernel": "attention fwd / bwd", "case": label, "B,H,d,Nq,Nkv": [B, H, d, Nq, Nkv], "fwd_ms": ms_f,
                 "bwd_ms": ms_b, "fwd_TFLOPs": fl / ms_f / 1e9, "bwd_TFLOPs": 2.5 * fl / ms_b / 1e9})
    print(rows[-1], flush=True)
print(json.dumps({"rows": rows}))
