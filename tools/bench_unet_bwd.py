"""Timing of SDXL student forward+backward (LoRA) on the B200 engine: python tools/bench_unet_bwd.py [B]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.models.lora import LoraConfig
from flash.models.unets import DiffusersUNet2DCondWrapper
from oracle.unet import SDXL_KWARGS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
with torch.device("meta"):
    m = DiffusersUNet2DCondWrapper(**SDXL_KWARGS)
m = m.to_empty(device="cuda")
for p in m.parameters():
    torch.nn.init.normal_(p, std=0.02)
m.add_adapter(LoraConfig(r=64, lora_alpha=64, init_lora_weights="gaussian", target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
x = torch.randn(B, 4, 128, 128, device="cuda")
t = torch.full((B,), 500.0, device="cuda")
cond = {"cond": {"crossattn": torch.randn(B, 77, 2048, device="cuda"), "vector": torch.randn(B, 2816, device="cuda")}}
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    out = m(x, t, cond)
    torch.cuda.synchronize(); t1 = time.time()
    out.square().mean().backward()
    torch.cuda.synchronize(); t2 = time.time()
    print(f"iter {it}: fwd {1e3 * (t1 - t0):.1f} ms  bwd {1e3 * (t2 - t1):.1f} ms  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
