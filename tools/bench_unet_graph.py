"""SDXL teacher evaluation replayed from a CUDA graph (the form it has inside the step): python tools/bench_unet_graph.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.b200.graphs import GraphedDenoiser
from flash.models.unets import DiffusersUNet2DCondWrapper
from oracle.unet import SDXL_KWARGS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
with torch.device("meta"):
    m = DiffusersUNet2DCondWrapper(**SDXL_KWARGS)
m = m.to_empty(device="cuda")
for p in m.parameters():
    torch.nn.init.normal_(p, std=0.02)
m.freeze()
x = torch.randn(B, 4, 128, 128, device="cuda")
t = torch.full((B,), 500.0, device="cuda")
cond = {"cond": {"crossattn": torch.randn(B, 77, 2048, device="cuda"), "vector": torch.randn(B, 2816, device="cuda")}}
g = GraphedDenoiser(m)
with torch.no_grad():
    ref = m(x, t, cond).clone()
    for _ in range(3):
        out = g(x, t, cond)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        out = g(x, t, cond, clone=False)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
rel = ((out - ref).norm() / ref.norm()).item()
print(f"SDXL UNet fwd B={B} (graph replay, FD_PDL={os.environ.get('FD_PDL', 'default')}): {ms:.2f} ms -> "
      f"{6.76 * B / ms * 1e3:.0f} TFLOP/s; replay vs eager rel diff {rel:.2e}")
