"""LPIPS (net="vgg") on the B200 kernels — the default `distill_loss_type="lpips"` of every example yaml
(reference src/flash/models/flash/flash_diffusion_model.py:102-103 `lpips.LPIPS(net="vgg")`, :383-397).

Same call contract as the `lpips` package (`LPIPS(net="vgg")(in0, in1) -> [B, 1, 1, 1]`, inputs in [-1, 1]) and the
same state-dict keys (`net.slice1.0.weight`, `lin0.model.1.weight`, ...), so the published weights load; offline they
are random.  The 13 VGG16 convolutions run as implicit GEMMs with the ReLU in the epilogue (fd_gemm conv mode, act 2),
pooling / unit-normalised squared distance / their gradients in small NHWC kernels (fd_maxpool2x2, fd_lpips_layer).
Gradients flow to `in0` only (the student image; `in1`, the teacher image, is a constant of the objective); the VGG
weights are frozen as in the reference.  Math restated in oracle/lpips.py.  No CPU fallback.
"""
import torch
import torch.nn as nn

from ..b200 import ops, raw
from ..b200.ops import ConvPack

VGG_SLICES = [[(0, 3, 64), (2, 64, 64)],
              [(5, 64, 128), (7, 128, 128)],
              [(10, 128, 256), (12, 256, 256), (14, 256, 256)],
              [(17, 256, 512), (19, 512, 512), (21, 512, 512)],
              [(24, 512, 512), (26, 512, 512), (28, 512, 512)]]
CHNS = [64, 128, 256, 512, 512]


class _Slice(nn.Module):
    def __init__(self, convs):
        super().__init__()
        self.idx = [i for i, _, _ in convs]
        for i, cin, cout in convs:
            self.add_module(str(i), nn.Conv2d(cin, cout, 3, padding=1))


class _VGG(nn.Module):
    def __init__(self):
        super().__init__()
        for k, convs in enumerate(VGG_SLICES):
            setattr(self, f"slice{k + 1}", _Slice(convs))


class _Lin(nn.Module):
    def __init__(self, chn):
        super().__init__()
        self.model = nn.Sequential(nn.Dropout(), nn.Conv2d(chn, 1, 1, bias=False))


class _ConvReluFn(torch.autograd.Function):
    """y = relu(conv3x3(x) + b) with the ReLU in the GEMM epilogue; backward = ReLU mask, then the data-gradient conv."""

    @staticmethod
    def forward(ctx, x, pack, geom):
        p = pack.pack()
        y = ops._conv_fwd_raw(x, geom, p["w"], p["b"], None, None, None, 1, pack.cin, act=2)
        ctx.pack, ctx.geom = pack, geom
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.conv_dgrad(raw.relu_bwd(y, dy.contiguous()), ctx.geom, ctx.pack, 1), None, None


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, geom):
        NB, H, W = geom
        ctx.geom = geom
        ctx.save_for_backward(x)
        return raw.maxpool2x2(x, NB, H, W, x.shape[1])

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        NB, H, W = ctx.geom
        return raw.maxpool2x2_bwd(x, dy.contiguous(), NB, H, W, x.shape[1]), None


class _LayerDistFn(torch.autograd.Function):
    """out[n] += distance of one feature layer; gradient to f0 only."""

    @staticmethod
    def forward(ctx, f0, f1, w, acc, NB, HW):
        C = f0.shape[1]
        ctx.meta = (NB, HW, C)
        ctx.save_for_backward(f0, f1, w)
        out = acc.clone()
        raw.lpips_layer(f0, f1, w, out, NB, HW, C)
        return out

    @staticmethod
    def backward(ctx, gout):
        f0, f1, w = ctx.saved_tensors
        NB, HW, C = ctx.meta
        return raw.lpips_layer_bwd(f0, f1, w, gout.float().contiguous(), NB, HW, C), None, None, gout, None, None


class LPIPS(nn.Module):
    def __init__(self, net: str = "vgg", **unused):
        super().__init__()
        if net != "vgg":
            raise NotImplementedError("only net='vgg' (the reference's choice) is built")
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, :, None, None])
        self.net = _VGG()
        for k, c in enumerate(CHNS):
            setattr(self, f"lin{k}", _Lin(c))
            with torch.no_grad():
                getattr(self, f"lin{k}").model[1].weight.abs_()           # lpips' lin weights are non-negative
        for p in self.parameters():
            p.requires_grad = False
        self.eval()
        self.__dict__["_packs"] = {}

    def _pack(self, key, make):
        packs = self.__dict__.setdefault("_packs", {})
        if key not in packs:
            packs[key] = make()
        return packs[key]

    def _features(self, img):
        """img [B, 3, H, W] fp32 in [-1, 1] -> the five tap tensors (NHWC bf16 rows) and their geometries."""
        NB, _, H, W = img.shape
        x = ops.to_nhwc((img - self.shift) / self.scale, 8)
        geom = (NB, H, W)
        taps = []
        for k in range(5):
            sl = getattr(self.net, f"slice{k + 1}")
            if k > 0:
                x = _MaxPoolFn.apply(x, geom) if torch.is_grad_enabled() and x.requires_grad else \
                    raw.maxpool2x2(x, geom[0], geom[1], geom[2], x.shape[1])
                geom = (NB, geom[1] // 2, geom[2] // 2)
            for i in sl.idx:
                conv = getattr(sl, str(i))
                pack = self._pack(("c", id(conv)), lambda: ConvPack(conv))
                if torch.is_grad_enabled() and x.requires_grad:
                    x = _ConvReluFn.apply(x, pack, geom)
                else:
                    p = pack.pack()
                    x = ops._conv_fwd_raw(x, geom, p["w"], p["b"], None, None, None, 1, pack.cin, act=2)
            taps.append((x, geom))
        return taps

    def forward(self, in0, in1, normalize: bool = False):
        if not in0.is_cuda:
            raise RuntimeError("LPIPS runs only on CUDA (B200) tensors: there is no CPU fallback")
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        if in0.shape[2] % 16 or in0.shape[3] % 16:
            raise ValueError("image size must be a multiple of 16 (four 2x2 poolings)")
        with torch.no_grad():
            t1 = self._features(in1.detach().float())
        t0 = self._features(in0.float())
        NB = in0.shape[0]
        acc = torch.zeros(NB, device=in0.device, dtype=torch.float32)
        for k in range(5):
            (f0, geom), (f1, _) = t0[k], t1[k]
            w = self._pack(("w", k), lambda: getattr(self, f"lin{k}").model[1].weight.detach().float().reshape(-1).contiguous())
            acc = _LayerDistFn.apply(f0, f1, w, acc, NB, geom[1] * geom[2])
        return acc.view(NB, 1, 1, 1)
