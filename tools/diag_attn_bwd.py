"""Diagnostic: accuracy of the attention backward on COMMON-MODE-dominated inputs (keys / values = large mean + small
per-token variation, as in deep blocks of a random-init UNet), ours vs torch SDPA (bf16) vs an fp64 reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch
import torch.nn.functional as F

from flash.b200 import raw


def cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float(torch.dot(a, b) / (a.norm() * b.norm()))


torch.manual_seed(0)
B, H, N, d = 1, 20, 1024, 64
for cm in [0.0, 1.0, 4.0, 16.0]:
    mk = lambda: (cm * torch.randn(1, 1, H * d, device="cuda") + torch.randn(B, N, H * d, device="cuda")).bfloat16()
    q, k, v = mk(), mk(), mk()
    do = (torch.randn(B, N, H * d, device="cuda") * 0.1).bfloat16()
    # fp64 reference
    q4, k4, v4 = (t.double().view(B, N, H, d).transpose(1, 2).requires_grad_(True) for t in (q, k, v))
    s = (q4 @ k4.transpose(-1, -2)) / d ** 0.5
    o_ref = torch.softmax(s, -1) @ v4
    gq, gk, gv = torch.autograd.grad(o_ref, (q4, k4, v4), do.double().view(B, N, H, d).transpose(1, 2))
    # ours
    o, lse = raw.attention_fwd(q, k, v, H, need_lse=True)
    dq, dk, dv = raw.attention_bwd(q, k, v, o, lse, do, H)
    f = lambda t: t.view(B, N, H, d).transpose(1, 2)
    # SDPA bf16
    qb, kb, vb = (f(t).detach().clone().requires_grad_(True) for t in (q, k, v))
    ob = F.scaled_dot_product_attention(qb, kb, vb)
    sq, sk, sv = torch.autograd.grad(ob, (qb, kb, vb), f(do))
    print(f"common-mode {cm:5.1f}: fwd rel ours {float((f(o).double() - o_ref).norm() / o_ref.norm()):.2e} "
          f"sdpa {float((ob.double() - o_ref).norm() / o_ref.norm()):.2e} | cos(dq) ours {cos(f(dq), gq):.5f} sdpa {cos(sq, gq):.5f} | "
          f"cos(dk) ours {cos(f(dk), gk):.5f} sdpa {cos(sk, gk):.5f} | cos(dv) ours {cos(f(dv), gv):.5f} sdpa {cos(sv, gv):.5f}")
