"""Checker-side diagnostic (uses oracle/, like the tests; never imported by the product or bench.py): decoder input-gradient agreement (product vs fp32 oracle) as a function of the latent channel count."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch
from flash.models.vae import AutoencoderKL
from oracle.vae import AutoencoderKLOracle

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cos = lambda a, b: (torch.dot(a.float().reshape(-1), b.float().reshape(-1)) / (a.float().norm() * b.float().norm() + 1e-30)).item()
rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
for lc, quant, hw, seed in [(4, True, 8, 1), (4, False, 8, 1), (8, False, 8, 2), (16, False, 8, 2), (16, True, 8, 2),
                            (16, False, 16, 3), (32, False, 8, 2)]:
    kw = dict(latent_channels=lc, use_quant_conv=quant, use_post_quant_conv=quant)
    torch.manual_seed(seed)
    ora = AutoencoderKLOracle(**kw).cuda()
    with torch.no_grad():
        for n, p in ora.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    prod = AutoencoderKL(**kw).cuda()
    prod.load_state_dict(ora.state_dict())
    z = torch.randn(2, lc, hw, hw, device="cuda")
    zp, zo = z.clone().requires_grad_(True), z.clone().requires_grad_(True)
    out, ref = prod.decode(zp), ora.decoder(ora.post_quant_conv(zo))
    g = torch.randn_like(ref)
    (out * g).sum().backward(); (ref * g).sum().backward()
    per_c = [round(cos(zp.grad[:, c], zo.grad[:, c]), 4) for c in range(lc)]
    # the same through a bf16-autocast oracle: the noise floor of bf16 arithmetic on this graph
    zb = z.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        refb = ora.decoder(ora.post_quant_conv(zb))
    (refb.float() * g).sum().backward()
    print(f"latent={lc} quant={quant} hw={hw}: fwd rel {rel(out, ref):.2e}  grad cos {cos(zp.grad, zo.grad):.5f} rel {rel(zp.grad, zo.grad):.2e}"
          f" | bf16-autocast oracle: cos {cos(zb.grad, zo.grad):.5f} rel {rel(zb.grad, zo.grad):.2e} | per-channel min {min(per_c)} {per_c if lc <= 16 else ''}", flush=True)
