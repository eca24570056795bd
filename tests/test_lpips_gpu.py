"""GPU parity of the LPIPS-VGG distillation loss (reference flash_diffusion_model.py:102-103,383-397)
against the fp32 oracle (oracle/lpips.py restating lpips==0.1.4, oracle/vae.py), random weights on both sides."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _cos(a, b):
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def test_maxpool_and_relu_kernels():
    from flash.b200 import raw
    torch.manual_seed(0)
    NB, H, W, C = 2, 16, 24, 64
    x = torch.randn(NB, H, W, C, device="cuda").bfloat16()
    y = raw.maxpool2x2(x.view(-1, C), NB, H, W, C).view(NB, H // 2, W // 2, C)
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(xf, 2, 2)
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1))
    dy = torch.randn(NB, H // 2, W // 2, C, device="cuda").bfloat16()
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dx = raw.maxpool2x2_bwd(x.view(-1, C), dy.view(-1, C), NB, H, W, C).view(NB, H, W, C)
    assert torch.equal(dx.float(), xf.grad.permute(0, 2, 3, 1))
    r = torch.relu(torch.randn(4096, 64, device="cuda")).bfloat16()
    g = torch.randn(4096, 64, device="cuda").bfloat16()
    assert torch.equal(raw.relu_bwd(r, g), torch.where(r > 0, g, torch.zeros_like(g)))


@pytest.mark.parametrize("C,HW", [(64, 4096), (512, 256), (256, 1000)])
def test_lpips_layer_kernels(C, HW):
    from flash.b200 import raw
    torch.manual_seed(C)
    NB = 3
    f0 = torch.relu(torch.randn(NB * HW, C, device="cuda")).bfloat16()
    f1 = torch.relu(torch.randn(NB * HW, C, device="cuda")).bfloat16()
    w = torch.rand(C, device="cuda")
    out = torch.full((NB,), 0.5, device="cuda")
    raw.lpips_layer(f0, f1, w, out, NB, HW, C)
    a = f0.float().view(NB, HW, C).requires_grad_(True)
    b = f1.float().view(NB, HW, C)
    u = a / (a.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    g = b / (b.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    ref = ((u - g) ** 2 * w).sum(-1).mean(-1)
    assert torch.allclose(out - 0.5, ref, rtol=2e-3, atol=1e-5)
    gout = torch.randn(NB, device="cuda")
    (ref * gout).sum().backward()
    df0 = raw.lpips_layer_bwd(f0, f1, w, gout, NB, HW, C)
    assert _rel(df0.view(NB, HW, C), a.grad) < 1e-2


def _pair():
    from flash.models.lpips import LPIPS
    from oracle.lpips import LPIPSOracle
    torch.manual_seed(0)
    ora = LPIPSOracle().cuda()
    prod = LPIPS(net="vgg").cuda()
    prod.load_state_dict(ora.state_dict())
    return prod, ora


@pytest.mark.parametrize("B,px", [(2, 64), (1, 256)])
def test_lpips_forward_and_gradient(B, px):
    prod, ora = _pair()
    x = (torch.rand(B, 3, px, px, device="cuda") * 2 - 1)
    y = (x + 0.3 * torch.randn_like(x)).clamp(-1, 1)
    xp, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    out, ref = prod(xp, y), ora(xo, y)
    assert out.shape == ref.shape == (B, 1, 1, 1)
    assert _rel(out, ref) < 2e-2, (_rel(out, ref), out.flatten(), ref.flatten())
    out.mean().backward()
    ref.mean().backward()
    assert _cos(xp.grad, xo.grad) > 0.99, _cos(xp.grad, xo.grad)
    with torch.no_grad():
        assert float(prod(x, x).abs().max()) < 1e-6


def test_lpips_distill_loss_through_the_vae_decoder():
    """the whole reference branch (:383-397): crop 64x64 latents, decode student and teacher, clamp, LPIPS, mean — value
    and gradient with respect to the student latents against the fp32 oracle."""
    from flash.models.vae import AutoencoderKL
    from oracle.lpips import lpips_distill_loss
    from oracle.vae import AutoencoderKLOracle
    prod_l, ora_l = _pair()
    torch.manual_seed(1)
    ora_v = AutoencoderKLOracle(scaling_factor=1.0).cuda()
    prod_v = AutoencoderKL(scaling_factor=1.0).cuda()
    prod_v.load_state_dict(ora_v.state_dict())
    for p in list(prod_v.parameters()) + list(ora_v.parameters()) + list(ora_l.parameters()):
        p.requires_grad = False
    s = torch.randn(1, 4, 96, 96, device="cuda")
    t = s + 0.2 * torch.randn_like(s)
    sp, so = s.clone().requires_grad_(True), s.clone().requires_grad_(True)

    class V:                      # the wrapper's decode contract without scaling
        def __init__(self, m):
            self.m = m

        def decode(self, z):
            return self.m.decode(z)

    loss_o = lpips_distill_loss(ora_l, V(ora_v), so, t)
    ch = (96 - 64) // 2
    dec_s = prod_v.decode(sp[:, :, ch:ch + 64, ch:ch + 64]).clamp(-1, 1)
    with torch.no_grad():
        dec_t = prod_v.decode(t[:, :, ch:ch + 64, ch:ch + 64]).clamp(-1, 1)
    loss_p = prod_l(dec_s, dec_t).mean()
    assert abs(float(loss_p) - float(loss_o)) / abs(float(loss_o)) < 3e-2, (float(loss_p), float(loss_o))
    loss_p.backward()
    loss_o.backward()
    assert _cos(sp.grad, so.grad) > 0.98, _cos(sp.grad, so.grad)
    assert float(sp.grad[:, :, :ch].abs().max()) == 0          # outside the crop: no gradient
