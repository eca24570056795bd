"""Quick timing of one B200 UNet forward (no grad): python tools/bench_unet.py [B]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch

from flash.models.unets import DiffusersUNet2DCondWrapper
from oracle.unet import SDXL_KWARGS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
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
with torch.no_grad():
    for _ in range(2):
        m(x, t, cond)
    torch.cuda.synchronize()
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        m(x, t, cond)
    issue = (time.time() - t0) / n * 1e3
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    wall = (time.time() - t0) / n * 1e3
print(f"SDXL UNet fwd B={B}: {ms:.2f} ms GPU ({wall:.2f} ms wall, host issue {issue:.2f} ms)  -> "
      f"{6.76 * B / ms * 1e3:.0f} TFLOP/s ({6.76 * B / ms / 1374.2 * 1000 * 100:.1f}% of sustained bf16 peak)")
from flash.b200 import lib as fdlib
import ctypes
l = fdlib.load()
l.fd_profile_enable(1)
with torch.no_grad():
    m(x, t, cond)
l.fd_profile_enable(0)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
l.fd_profile_dump(os.path.join(ROOT, 'gpurun_out', f'gemm_shapes_B{B}.csv').encode())
ms_, fl_, cnt_ = (ctypes.c_double * 4)(), (ctypes.c_double * 4)(), (ctypes.c_longlong * 4)()
l.fd_profile_summary(ms_, fl_, cnt_, 4)
for i, name in enumerate(["gemm", "conv", "attn_fwd", "attn_bwd"]):
    if cnt_[i]:
        print(f"  {name:9s} launches {cnt_[i]:5d}  {ms_[i]:8.2f} ms  {fl_[i] / ms_[i] / 1e9:8.1f} TFLOP/s")
