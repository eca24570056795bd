"""One attention backward launch per shape, for `ncu -k regex:attn_bwd` captures: python tools/prof_attn_bwd.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.b200 import raw

for (B, H, N) in [(4, 20, 1024), (2, 24, 4250)]:
    q, k, v, do = (torch.randn(B, N, H * 64, device="cuda").bfloat16() for _ in range(4))
    o, lse = raw.attention_fwd(q, k, v, H, need_lse=True)
    for _ in range(2):
        raw.attention_bwd(q, k, v, o, lse, do, H)
    torch.cuda.synchronize()
