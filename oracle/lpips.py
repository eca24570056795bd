"""ORACLE (test infrastructure — never imported by the product path).

fp32 restatement of `lpips.LPIPS(net="vgg")` (lpips==0.1.4, reference setup.py:40; used at reference
src/flash/models/flash/flash_diffusion_model.py:102-103 and :383-397) and of the reference's lpips distillation loss:
center-crop both latents to 64x64, decode both with the VAE, clamp to [-1, 1], LPIPS, mean.

Published algorithm (Zhang et al. 2018, lpips v0.1): ScalingLayer ((x - shift) / scale), VGG16 feature taps relu1_2,
relu2_2, relu3_3, relu4_3, relu5_3 (64, 128, 256, 512, 512 channels), unit-normalise every pixel's feature vector
(x / (||x||_2 + 1e-10)), squared difference, per-layer non-negative 1x1 "lin" weights, spatial mean, sum over layers.
State-dict keys as in the lpips package (`net.slice1.0.weight`, `lin0.model.1.weight`, ...) so real weights would load.

PARITY: the VGG16 feature stack (13 convolutions, pooling, the five taps) is pinned to torchvision's own
`vgg16().features` — the module the lpips package wraps — in tests/test_lpips_vgg_torchvision_cpu.py.  UNPINNED: the
lpips-specific glue (scaling layer, unit normalisation, lin layers, reduction), restated from the paper / package source;
neither the lpips package nor the trained VGG16 / lin weights are available offline (random weights throughout).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

# torchvision vgg16.features indices of the 13 convolutions, grouped by lpips' five slices
VGG_SLICES = [[(0, 3, 64), (2, 64, 64)],
              [(5, 64, 128), (7, 128, 128)],
              [(10, 128, 256), (12, 256, 256), (14, 256, 256)],
              [(17, 256, 512), (19, 512, 512), (21, 512, 512)],
              [(24, 512, 512), (26, 512, 512), (28, 512, 512)]]
CHNS = [64, 128, 256, 512, 512]


class _Slice(nn.Module):
    def __init__(self, convs, pool_first):
        super().__init__()
        self.pool_first = pool_first
        self.idx = [i for i, _, _ in convs]
        for i, cin, cout in convs:
            self.add_module(str(i), nn.Conv2d(cin, cout, 3, padding=1))

    def forward(self, x):
        if self.pool_first:
            x = F.max_pool2d(x, 2, 2)
        for i in self.idx:
            x = F.relu(getattr(self, str(i))(x))
        return x


class _VGG(nn.Module):
    def __init__(self):
        super().__init__()
        for k, convs in enumerate(VGG_SLICES):
            setattr(self, f"slice{k + 1}", _Slice(convs, pool_first=k > 0))

    def forward(self, x):
        outs = []
        for k in range(5):
            x = getattr(self, f"slice{k + 1}")(x)
            outs.append(x)
        return outs


class _Lin(nn.Module):
    def __init__(self, chn):
        super().__init__()
        self.model = nn.Sequential(nn.Dropout(), nn.Conv2d(chn, 1, 1, bias=False))


class LPIPSOracle(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, :, None, None])
        self.net = _VGG()
        for k, c in enumerate(CHNS):
            setattr(self, f"lin{k}", _Lin(c))
            with torch.no_grad():                    # lpips' lin weights are non-negative
                getattr(self, f"lin{k}").model[1].weight.abs_()
        self.eval()

    @staticmethod
    def _unit(x, eps=1e-10):
        return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)

    def forward(self, in0, in1):
        f0 = self.net((in0 - self.shift) / self.scale)
        f1 = self.net((in1 - self.shift) / self.scale)
        val = 0
        for k in range(5):
            d = (self._unit(f0[k]) - self._unit(f1[k])) ** 2
            val = val + getattr(self, f"lin{k}").model[1](d).mean([2, 3], keepdim=True)
        return val


def lpips_distill_loss(lpips, vae, student_output, teacher_output):
    """reference flash_diffusion_model.py:383-397"""
    ch = (student_output.shape[2] - 64) // 2
    cw = (student_output.shape[3] - 64) // 2
    s = student_output[:, :, ch:ch + 64, cw:cw + 64]
    t = teacher_output[:, :, ch:ch + 64, cw:cw + 64]
    return lpips(vae.decode(s).clamp(-1, 1), vae.decode(t).clamp(-1, 1)).mean()
