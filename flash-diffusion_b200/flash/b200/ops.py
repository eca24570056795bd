"""Autograd-aware operators over the raw C-ABI kernels.

Every operator here runs hand-written sm_100a kernels (flash.b200.raw) in both directions.  Under
`torch.no_grad()` (frozen teacher, ~85% of the step) the raw kernel is called directly; when a
gradient is required a `torch.autograd.Function` records the tensors its hand-written backward needs.

Only activation gradients (dX) and LoRA A/B gradients exist: every other weight is frozen in the
Flash-Diffusion step (reference: examples/train_flash_sdxl.py:206-219, src/flash/trainer/trainer.py:115-124).

Activations are channels-last bf16 matrices [NB*H*W, C]; `geom` = (NB, H, W).
"""
import os

import torch

from . import raw

BF16 = torch.bfloat16


def _grad_on(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _version(p):
    return (p.data_ptr(), p._version, p.device)


class PackCache:
    """Per-module cache of kernel-ready (bf16, re-laid-out) copies of fp32 parameters, keyed on the
    parameters' version counters so that optimizer updates (LoRA) invalidate them."""

    def __init__(self):
        self.store = {}

    def get(self, key, params, build):
        sig = tuple(_version(p) for p in params)
        hit = self.store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self.store[key] = (sig, val)
        return val


def cache_of(module) -> PackCache:
    c = module.__dict__.get("_fd_cache")
    if c is None:
        c = PackCache()
        module.__dict__["_fd_cache"] = c
    return c


# ------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------
def pack_matrix(w, scale=1.0):
    """fp32 [N, K] -> bf16 [N, K] (K padded to a multiple of 8 is the caller's business)."""
    return raw.cast_scale(w.detach().reshape(w.shape[0], -1).float().contiguous(), scale)


def pack_conv3x3(w):
    """[Cout, Cin, kh, kw] -> bf16 [Cout, taps * ceil64(Cin)], K ordered (tap, channel)."""
    Cout, Cin, kh, kw = w.shape
    cpad = (Cin + 63) // 64 * 64
    buf = torch.zeros((Cout, kh * kw, cpad), device=w.device, dtype=torch.float32)
    buf[:, :, :Cin] = w.detach().float().permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin)
    return raw.cast_scale(buf.reshape(Cout, kh * kw * cpad), 1.0)


def pack_conv3x3_dgrad(w):
    """weights of the data-gradient convolution: Wd[ci, (kh',kw'), co] = W[co, ci, 2-kh', 2-kw']."""
    wd = w.detach().float().flip(2, 3).permute(1, 0, 2, 3).contiguous()   # [Cin, Cout, kh', kw']
    return pack_conv3x3(wd)


def pack_geglu(w, b):
    """diffusers GEGLU.proj [2*inner, C] = [value | gate] -> 16-row interleaved blocks (fd_gemm geglu)."""
    two_inner, C = w.shape
    inner = two_inner // 2
    wv, wg = w.detach().float()[:inner].reshape(-1, 16, C), w.detach().float()[inner:].reshape(-1, 16, C)
    wi = torch.stack([wv, wg], dim=1).reshape(two_inner, C).contiguous()
    bi = torch.stack([b.detach().float()[:inner].reshape(-1, 16), b.detach().float()[inner:].reshape(-1, 16)],
                     dim=1).reshape(-1).contiguous()
    return raw.cast_scale(wi, 1.0), bi


def _f32(p):
    return None if p is None else p.detach().float().contiguous()


# ------------------------------------------------------------------------------------------------
# convolution (implicit GEMM)
# ------------------------------------------------------------------------------------------------
# taps of the stride-2 3x3 conv (pad 1) over the space-to-depth input [4*NB, H/2, W/2, C]:
# input row 2*ho + kh - 1 -> phase (kh-1)&1, row offset -1 for kh = 0 else 0
def _s2_taps(NB):
    taps = []
    for kh in range(3):
        for kw in range(3):
            ph, pw = (kh - 1) & 1, (kw - 1) & 1
            dh, dw = (-1 if kh == 0 else 0), (-1 if kw == 0 else 0)
            taps.append(((ph * 2 + pw) * NB, dh, dw))
    return taps


def _s2_taps_asym(NB):
    """taps of diffusers' Downsample2D(padding=0) (the VAE encoder): F.pad(x, (0, 1, 0, 1)) then a 3x3 stride-2 conv
    without padding — input row 2*ho + kh: phase kh & 1, row offset kh >> 1 (the TMA zero-fill supplies the pad)."""
    return [(((kh & 1) * 2 + (kw & 1)) * NB, kh >> 1, kw >> 1) for kh in range(3) for kw in range(3)]


def _conv_fwd_raw(x, geom, wpack, bias, rowvec, residual, shortcut, stride, Cin, out_fp32=False, pad_mode="same",
                  act=0, colstats=None):
    NB, H, W = geom
    if stride == 1:
        conv = dict(NB_in=NB, H=H, W=W, C=Cin, taps=raw.TAPS_3X3)
        a1, M, rpg = x, NB * H * W, H * W
    else:
        a1 = raw.space_to_depth(x, NB, H, W, Cin)
        conv = dict(NB_in=4 * NB, H=H // 2, W=W // 2, C=Cin,
                    taps=_s2_taps(NB) if pad_mode == "same" else _s2_taps_asym(NB))
        M, rpg = NB * (H // 2) * (W // 2), (H // 2) * (W // 2)
    a2 = b2 = None
    if shortcut is not None:
        a2, b2 = shortcut
    return raw.gemm(a1, wpack, a2=a2, b2=b2, bias=bias, rowvec=rowvec, rows_per_group=rpg, residual=residual,
                    conv=conv, M=M, out_fp32=out_fp32, act=act, colstats=colstats)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, x2, mod, geom, stride, rowvec, out_fp32):
        p = mod.pack()
        shortcut = None
        if x2 is not None:
            shortcut = (x2, mod.pack_shortcut())
        y = _conv_fwd_raw(x, geom, p["w"], p["b"], rowvec, residual, shortcut, stride, mod.cin, out_fp32)
        ctx.mod, ctx.geom, ctx.stride = mod, geom, stride
        ctx.has_res, ctx.has_x2 = residual is not None, x2 is not None
        ctx.out_fp32 = out_fp32
        return y

    @staticmethod
    def backward(ctx, dy):
        mod, (NB, H, W), stride = ctx.mod, ctx.geom, ctx.stride
        if dy.dtype != BF16:
            dy = raw.cast_scale(dy.contiguous().float(), 1.0)
        dy = dy.contiguous()
        dx = dres = dx2 = None
        if ctx.needs_input_grad[0]:
            dx = conv_dgrad(dy, (NB, H, W), mod, stride)
        if ctx.has_res and ctx.needs_input_grad[1]:
            dres = dy
        if ctx.has_x2 and ctx.needs_input_grad[2]:
            dx2 = raw.gemm(dy, mod.pack_shortcut_t())
        return dx, dres, dx2, None, None, None, None, None


def conv_dgrad(dy, geom, mod, stride):
    """dX of a 3x3 conv.  geom is the INPUT geometry (NB, H, W)."""
    NB, H, W = geom
    cout = mod.cout
    if stride == 1:
        wd = mod.pack_dgrad()
        return raw.gemm(dy, wd, conv=dict(NB_in=NB, H=H, W=W, C=cout, taps=raw.TAPS_3X3), M=NB * H * W)
    # stride 2: per input phase (p, q) a small conv over dy (grid H/2 x W/2); results are written
    # phase-major and scattered back with depth_to_space.
    Ho, Wo = H // 2, W // 2
    Mo = NB * Ho * Wo
    out = torch.empty((4 * Mo, mod.cin), device=dy.device, dtype=BF16)
    packs = mod.pack_dgrad_s2()
    for ph in range(4):
        taps, wd = packs[ph]
        raw.gemm(dy, wd, conv=dict(NB_in=NB, H=Ho, W=Wo, C=cout, taps=taps), M=Mo, out=out[ph * Mo:(ph + 1) * Mo])
    return raw.depth_to_space(out, NB, H, W, mod.cin)


_NO_COLSTATS = (os.environ.get("FD_NO_COLSTATS") is not None
                or os.environ.get("FD_NO_TMA_STORE") is not None)     # the statistics live in the TMA-store epilogue


def colstats_of(t):
    """Per-image column sums [NB, C, 2] the producing GEMM / conv attached to its output (None if it did not)."""
    return getattr(t, "_fd_colstats", None)


def _take_colstats(arena, NB, rows_per_image, N, out_fp32=False):
    if arena is None or _NO_COLSTATS or out_fp32 or N % 32 or rows_per_image % 32:
        return None
    return arena.take(NB * N).view(NB, N, 2)


def conv3x3(x, geom, mod, *, rowvec=None, residual=None, x2=None, stride=1, out_fp32=False, pad_mode="same", act=0,
            arena=None):
    """y = conv3x3(x) + bias (+ rowvec per image) (+ residual) (+ x2 @ W_shortcut^T).  `mod` is a ConvPack.
    pad_mode="asym" (stride 2 only): the VAE encoder's Downsample2D(padding=0).  act=2: ReLU epilogue (no-grad only).
    arena (no-grad only): the epilogue also leaves the per-image column sums of y for a following GroupNorm
    (`colstats_of(y)`, consumed by `group_norm`)."""
    if _grad_on(x, residual, x2):
        if pad_mode != "same" or act:
            raise NotImplementedError("gradients through the asymmetric stride-2 conv / fused ReLU are not needed by the "
                                      "hot path (the VAE encoder is frozen and runs without a graph)")
        return _ConvFn.apply(x, residual, x2, mod, geom, stride, rowvec, out_fp32)
    p = mod.pack()
    shortcut = (x2, mod.pack_shortcut()) if x2 is not None else None
    NB, H, W = geom
    cs = _take_colstats(arena, NB, (H // stride) * (W // stride), mod.cout, out_fp32)
    y = _conv_fwd_raw(x, geom, p["w"], p["b"], rowvec, residual, shortcut, stride, mod.cin, out_fp32, pad_mode, act,
                      colstats=cs)
    if cs is not None:
        y._fd_colstats = cs
    return y


class ConvPack:
    """Kernel-side view of a 3x3 nn.Conv2d (+ optional 1x1 shortcut conv accumulated as K-segment 2)."""

    def __init__(self, conv, shortcut=None):
        self.conv, self.shortcut = conv, shortcut
        self.cout_true, self.cin_true = conv.weight.shape[0], conv.weight.shape[1]
        self.cin = (self.cin_true + 7) // 8 * 8          # channel padding of the NHWC input (conv_in)
        self.cout = (self.cout_true + 7) // 8 * 8        # and of the output (conv_out): zero rows
        self.cache = cache_of(conv)

    def _w(self):
        w = self.conv.weight.detach()
        if self.cin != self.cin_true or self.cout != self.cout_true:
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, self.cin - self.cin_true, 0, self.cout - self.cout_true))
        return w

    def pack(self):
        def build():
            b = _f32(self.conv.bias)
            if self.shortcut is not None and self.shortcut.bias is not None:
                b = b + _f32(self.shortcut.bias)
            if b is not None and self.cout != self.cout_true:
                b = torch.nn.functional.pad(b, (0, self.cout - self.cout_true))
            return {"w": pack_conv3x3(self._w()), "b": b}
        params = [self.conv.weight, self.conv.bias] + ([self.shortcut.bias] if self.shortcut is not None else [])
        return self.cache.get("fwd", params, build)

    def pack_shortcut(self):
        return self.cache.get("sc", [self.shortcut.weight], lambda: pack_matrix(self.shortcut.weight))

    def pack_shortcut_t(self):
        return self.cache.get("sc_t", [self.shortcut.weight],
                              lambda: pack_matrix(self.shortcut.weight.detach().reshape(self.cout, -1).t()))

    def pack_dgrad(self):
        return self.cache.get("dgrad", [self.conv.weight], lambda: pack_conv3x3_dgrad(self._w()))

    def pack_dgrad_s2(self):
        """per input phase: (taps over dy, packed weights [Cin, ntaps*ceil64(Cout)])."""
        def build():
            w = self._w().detach().float()                     # [Cout, Cin, 3, 3]
            cpad = (self.cout + 63) // 64 * 64
            packs = []
            for p in range(2):
                for q in range(2):
                    khs = [1] if p == 0 else [0, 2]            # input row 2a+p = 2*ho + kh - 1
                    kws = [1] if q == 0 else [0, 2]
                    taps, mats = [], []
                    for kh in khs:
                        for kw in kws:
                            dh = (p + 1 - kh) // 2             # ho = a + dh
                            dw = (q + 1 - kw) // 2
                            taps.append((0, dh, dw))
                            m = torch.zeros((self.cin, cpad), device=w.device)
                            m[:, :self.cout] = w[:, :, kh, kw].t()
                            mats.append(m)
                    wd = torch.stack(mats, dim=1).reshape(self.cin, len(taps) * cpad)
                    packs.append((taps, raw.cast_scale(wd.contiguous(), 1.0)))
            return packs
        return self.cache.get("dgrad_s2", [self.conv.weight], build)


# ------------------------------------------------------------------------------------------------
# GroupNorm (+SiLU) and LayerNorm
# ------------------------------------------------------------------------------------------------
class _GroupNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, geom, G, eps, silu):
        NB, H, W = geom
        C = x.shape[1]
        y, stats = raw.groupnorm_fwd(x, gamma, beta, NB, H * W, C, G, eps, silu, want_stats=True)
        ctx.save_for_backward(x, stats, gamma, beta)
        ctx.meta = (NB, H * W, C, G, silu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, gamma, beta = ctx.saved_tensors
        NB, HW, C, G, silu = ctx.meta
        dx = raw.groupnorm_bwd(x, stats, gamma, beta, dy.contiguous(), NB, HW, C, G, silu)
        return dx, None, None, None, None, None, None


def group_norm(x, geom, norm, silu):
    """torch.nn.GroupNorm parameters `norm`; x [NB*H*W, C] bf16."""
    cache = cache_of(norm)
    gamma, beta = cache.get("gb", [norm.weight, norm.bias], lambda: (_f32(norm.weight), _f32(norm.bias)))
    if _grad_on(x):
        return _GroupNormFn.apply(x, gamma, beta, geom, norm.num_groups, norm.eps, silu)
    NB, H, W = geom
    C = x.shape[1]
    cs = colstats_of(x)
    if cs is not None:
        return raw.groupnorm_apply_cols(x, cs, gamma, beta, NB, H * W, C, norm.num_groups, norm.eps, silu)
    return raw.groupnorm_fwd(x, gamma, beta, NB, H * W, C, norm.num_groups, norm.eps, silu)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        y, stats = raw.layernorm_fwd(x, gamma, beta, eps, save_stats=True)
        ctx.save_for_backward(x, stats, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, gamma = ctx.saved_tensors
        return raw.layernorm_bwd(x, stats, gamma, dy.contiguous()), None, None, None


def layer_norm(x, norm):
    cache = cache_of(norm)
    gamma, beta = cache.get("gb", [norm.weight, norm.bias], lambda: (_f32(norm.weight), _f32(norm.bias)))
    if _grad_on(x):
        return _LayerNormFn.apply(x, gamma, beta, norm.eps)
    return raw.layernorm_fwd(x, gamma, beta, norm.eps)


# ------------------------------------------------------------------------------------------------
# Linear (+LoRA, +GEGLU, +residual)
# ------------------------------------------------------------------------------------------------
class LinearPack:
    """Kernel-side view of one or several nn.Linear (or LoRA-wrapped Linear) sharing an input; several
    are fused along N (self-attention q/k/v, cross-attention k/v)."""

    def __init__(self, layers, geglu=False, head_pad=None, pad_cols=False):
        """head_pad = (H, d, dp): zero-pad every head from d to dp channels — output rows of q/k/v projections
        (pad_cols=False) or input columns of the attention out-projection (pad_cols=True) — so that the attention
        kernel sees a head dim that is a multiple of 16 (softmax(QK^T)V is unchanged by zero channels)."""
        self.layers = layers if isinstance(layers, (list, tuple)) else [layers]
        self.geglu = geglu
        self.head_pad, self.pad_cols = head_pad, pad_cols
        self.bases = [getattr(l, "base_layer", l) for l in self.layers]
        self.loras = [l if hasattr(l, "base_layer") else None for l in self.layers]
        self.has_lora = any(l is not None for l in self.loras)
        self.cache = cache_of(self.bases[0])
        self.N = sum(b.weight.shape[0] for b in self.bases)
        self.K = self.bases[0].weight.shape[1]
        # TMA needs 16-byte row pitches: an input width that is not a multiple of 8 bf16 (a 123-wide text context, a
        # 12-wide class vector — the reference's own wrapper tests use such sizes) is zero-padded, in the packs'
        # K columns here and on the activation side in `pad_k`
        k_eff = self.K if not (head_pad is not None and pad_cols) else head_pad[0] * head_pad[2]
        self.k_pad = (k_eff + 7) // 8 * 8 if k_eff % 8 else 0
        if head_pad is not None and head_pad[1] != head_pad[2] and geglu:
            raise NotImplementedError("head padding with GEGLU packs")
        self.N_packed = None

    def _w2d(self, b):
        w = b.weight.detach().reshape(b.weight.shape[0], -1)          # Linear or 1x1 Conv2d
        if self.head_pad is not None and self.head_pad[1] != self.head_pad[2]:
            H, d, dp = self.head_pad
            if self.pad_cols:
                w = torch.nn.functional.pad(w.reshape(w.shape[0], H, d), (0, dp - d)).reshape(w.shape[0], H * dp)
            else:
                w = torch.nn.functional.pad(w.reshape(H, d, w.shape[1]), (0, 0, 0, dp - d)).reshape(H * dp, w.shape[1])
        if self.k_pad:
            w = torch.nn.functional.pad(w, (0, self.k_pad - w.shape[1]))
        return w

    def pad_k(self, x):
        """zero columns up to the packs' padded K (autograd-aware: the gradient of the padding is dropped)"""
        if self.k_pad and x.shape[1] != self.k_pad:
            x = torch.nn.functional.pad(x, (0, self.k_pad - x.shape[1]))
        return x

    def _bias(self, b):
        bias = b.bias
        if bias is None:
            return None
        bias = bias.detach().float()
        if self.head_pad is not None and self.head_pad[1] != self.head_pad[2] and not self.pad_cols:
            H, d, dp = self.head_pad
            bias = torch.nn.functional.pad(bias.reshape(H, d), (0, dp - d)).reshape(H * dp)
        return bias

    def pack(self):
        def build():
            ws = [self._w2d(b).float() for b in self.bases]
            bs = [self._bias(b) for b in self.bases]
            bias = None
            if any(x is not None for x in bs):
                bias = torch.cat([x if x is not None else
                                  torch.zeros(w.shape[0], device=w.device) for x, w in zip(bs, ws)]).contiguous()
            w = torch.cat(ws, dim=0)
            if self.geglu:
                wp, bp = pack_geglu(w, bias)
                return {"w": wp, "b": bp}
            return {"w": raw.cast_scale(w.contiguous(), 1.0), "b": bias}
        params = [b.weight for b in self.bases] + [b.bias for b in self.bases if b.bias is not None]
        return self.cache.get(("fwd", len(self.bases), self.geglu), params, build)

    def pack_t(self):
        """W^T [K, N] for the data gradient."""
        def build():
            w = torch.cat([self._w2d(b).float() for b in self.bases], dim=0)
            if self.geglu:
                w = pack_geglu(w, torch.zeros(w.shape[0], device=w.device))[0].float()
            return raw.cast_scale(w.t().contiguous(), 1.0)
        return self.cache.get(("t", len(self.bases), self.geglu), [b.weight for b in self.bases], build)

    def pack_ln(self, norm):
        """LayerNorm folded into this GEMM: B = W * gamma (bf16), colsum_n = sum_k bf16(W*gamma)[n,k] (what the tensor
        core really multiplies), bias' = b + W beta.  See fd_gemm (include/flashb200.h)."""
        def build():
            ws = [self._w2d(b).float() for b in self.bases]
            w = torch.cat(ws, dim=0)
            bs = [self._bias(b) for b in self.bases]
            bias = torch.cat([x if x is not None else torch.zeros(wi.shape[0], device=w.device)
                              for x, wi in zip(bs, ws)])
            gamma, beta = norm.weight.detach().float(), norm.bias.detach().float()
            bias = bias + w @ beta
            wg = w * gamma[None, :]
            if self.geglu:
                wp, bp = pack_geglu(wg, bias)
            else:
                wp, bp = raw.cast_scale(wg.contiguous(), 1.0), bias.contiguous()
            return {"w": wp, "b": bp, "colsum": wp.float().sum(dim=1).contiguous()}
        params = [b.weight for b in self.bases] + [b.bias for b in self.bases if b.bias is not None] + \
                 [norm.weight, norm.bias]
        return self.cache.get(("ln", len(self.bases), self.geglu), params, build)

    # LoRA: T = x [A_1;..;A_n]^T  (M x n*r);  y += T @ blockdiag(s B_i)^T
    def lora_params(self):
        ps = []
        for l in self.loras:
            if l is not None:
                ps += [l.lora_A["default"].weight, l.lora_B["default"].weight]
        return ps

    def pack_lora(self):
        def build():
            r = next(l.r for l in self.loras if l is not None)
            dev = self.bases[0].weight.device
            n = len(self.layers)
            padded = self.head_pad is not None and self.head_pad[1] != self.head_pad[2]
            a_rows, b_rows = [], []
            for i, (l, b) in enumerate(zip(self.loras, self.bases)):
                nout_true = b.weight.shape[0]
                if l is not None:
                    A = l.lora_A["default"].weight.detach().float()              # [r, K]
                    Bm = l.lora_B["default"].weight.detach().float() * l.scaling  # [Nout, r]
                else:
                    A = torch.zeros((r, b.weight.shape[1]), device=dev)
                    Bm = torch.zeros((nout_true, r), device=dev)
                if padded:
                    H, d, dp = self.head_pad
                    if self.pad_cols:       # out-projection: its INPUT columns (A's columns) are head-padded
                        A = torch.nn.functional.pad(A.reshape(r, H, d), (0, dp - d)).reshape(r, H * dp)
                    else:                   # q/k/v: OUTPUT rows (B's rows) are head-padded
                        Bm = torch.nn.functional.pad(Bm.reshape(H, d, r), (0, 0, 0, dp - d)).reshape(H * dp, r)
                if self.k_pad:
                    A = torch.nn.functional.pad(A, (0, self.k_pad - A.shape[1]))
                a_rows.append(A)
                b_rows.append(Bm)
            a_cat = torch.cat(a_rows, dim=0)                                         # [n*r, K']
            b_blk = torch.zeros((sum(x.shape[0] for x in b_rows), n * r), device=dev)
            row = 0
            for i, Bm in enumerate(b_rows):
                b_blk[row:row + Bm.shape[0], i * r:(i + 1) * r] = Bm
                row += Bm.shape[0]
            return {"a": raw.cast_scale(a_cat.contiguous(), 1.0), "b": raw.cast_scale(b_blk, 1.0),
                    "a_t": raw.cast_scale(a_cat.t().contiguous(), 1.0),
                    "b_t": raw.cast_scale(b_blk.t().contiguous(), 1.0), "r": r}
        return self.cache.get("lora", self.lora_params(), build)


def _linear_fwd_raw(x, pack: LinearPack, residual, out=None, act=0, out_fp32=False):
    p = pack.pack()
    a2 = b2 = None
    t = None
    if pack.has_lora:
        lp = pack.pack_lora()
        t = raw.gemm(x, lp["a"])
        a2, b2 = t, lp["b"]
    y = raw.gemm(x, p["w"], a2=a2, b2=b2, bias=p["b"], geglu=pack.geglu, residual=residual, out=out, act=act,
                 out_fp32=out_fp32)
    return y, t


class _LinearFn(torch.autograd.Function):
    """y = act(x W^T + b (+ LoRA)) (+ residual); GEGLU handled by _GegluFn.  act = 1 is the tanh-GELU of the GEMM
    epilogue: its backward recomputes the pre-activation (one extra GEMM) instead of storing it.  out_fp32 returns
    fp32 (the incoming gradient is rounded to bf16 for the tensor-core backward)."""

    @staticmethod
    def forward(ctx, x, residual, pack, act, out_fp32, *lora_params):
        y, t = _linear_fwd_raw(x, pack, residual, act=act, out_fp32=out_fp32)
        ctx.pack = pack
        ctx.act = act
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, t)
        return y

    @staticmethod
    def backward(ctx, dy):
        pack = ctx.pack
        x, t = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != BF16:
            dy = raw.cast_scale(dy.float(), 1.0)
        dres = dy if (ctx.has_res and ctx.needs_input_grad[1]) else None
        if ctx.act:
            p = pack.pack()
            lp = pack.pack_lora() if pack.has_lora else None
            pre = raw.gemm(x, p["w"], a2=t, b2=lp["b"] if lp else None, bias=p["b"])
            dy = raw.gelu_tanh_bwd(pre, dy)
        dx = None
        grads = [None] * len(pack.lora_params())
        if pack.has_lora:
            lp = pack.pack_lora()
            dt = raw.gemm(dy, lp["b_t"])                                  # [M, n*r] = dy @ (sB)
            if ctx.needs_input_grad[0]:
                dx = raw.gemm(dy, pack.pack_t(), a2=dt, b2=lp["a_t"])     # dy W + dt A
            grads = _lora_weight_grads(pack, lp, x, t, dy, dt)
        elif ctx.needs_input_grad[0]:
            dx = raw.gemm(dy, pack.pack_t())
        return (dx, dres, None, None, None, *grads)


def _lora_weight_grads(pack, lp, x, t, dy, dt):
    """dA_i = dt_i^T x ; dB_i = s * dy_i^T t_i   (fp32, in the parameters' layout).  With head-padded packs the zero
    channels the kernels see are dropped again: rows of dB for q/k/v projections, columns of dA for the attention
    out-projection."""
    r = lp["r"]
    x_t = raw.transpose(x)                      # [K, M]
    dy_t = raw.transpose(dy)                    # [N, M]
    t_t = raw.transpose(t)                      # [n*r, M]
    dt_t = raw.transpose(dt)                    # [n*r, M]
    d_a = raw.gemm(dt_t, x_t, out_fp32=True)    # [n*r, K]
    d_b = raw.gemm(dy_t, t_t, out_fp32=True)    # [N, n*r]  (block-diagonal part is what we keep)
    padded = pack.head_pad is not None and pack.head_pad[1] != pack.head_pad[2]
    grads = []
    row = 0
    for i, (l, b) in enumerate(zip(pack.loras, pack.bases)):
        nout = b.weight.shape[0]
        rows_here = nout
        if padded and not pack.pad_cols:
            H, d, dp = pack.head_pad
            rows_here = H * dp
        if l is not None:
            ga = d_a[i * r:(i + 1) * r]
            gb = d_b[row:row + rows_here, i * r:(i + 1) * r]
            if padded:
                H, d, dp = pack.head_pad
                if pack.pad_cols:
                    ga = ga.reshape(r, H, dp)[:, :, :d].reshape(r, H * d)
                else:
                    gb = gb.reshape(H, dp, r)[:, :d].reshape(H * d, r)
            if pack.k_pad:
                ga = ga[:, :l.lora_A["default"].weight.shape[1]]
            grads.append(ga)
            grads.append(gb * l.scaling)
        row += rows_here
    return grads


class StatsArena:
    """Row-statistics buffers of one denoiser evaluation carved out of a few large zero-filled blocks: one memset per
    block instead of one per producing GEMM (~210 per SDXL evaluation)."""

    def __init__(self, device, block_rows=1 << 21):
        self.device, self.block_rows = device, block_rows
        self.block, self.off = None, 0

    def take(self, M):
        if self.block is None or self.off + M > self.block.shape[0]:
            self.block = torch.zeros((max(self.block_rows, M), 2), device=self.device, dtype=torch.float32)
            self.off = 0
        out = self.block[self.off:self.off + M]
        self.off += M
        return out


def linear(x, pack: LinearPack, residual=None, want_stats=None, act=0, out_fp32=False, arena: "StatsArena" = None,
           colstats_images=0):
    """y = act(x W^T + b (+LoRA)) (+residual).  want_stats None: return y.  True/False: return (y, stats) where stats
    are the [M,2] row statistics of y (fused into the GEMM epilogue) when requested and the no-grad, LoRA-free path is
    taken, else None.  colstats_images = NB > 0 (no-grad, LoRA-free, with an arena): the epilogue leaves the per-image
    column sums of y for a following GroupNorm (`colstats_of(y)`)."""
    lora_params = pack.lora_params() if pack.has_lora else []
    x = pack.pad_k(x)
    if _grad_on(x, residual, *lora_params):
        y = _LinearFn.apply(x, residual, pack, act, out_fp32, *lora_params)
        return y if want_stats is None else (y, None)
    if colstats_images and not pack.has_lora and want_stats is None and not pack.geglu:
        p = pack.pack()
        N = p["w"].shape[0]
        cs = _take_colstats(arena, colstats_images, x.shape[0] // colstats_images, N, out_fp32)
        if cs is not None and x.shape[0] % colstats_images == 0:
            y = raw.gemm(x, p["w"], bias=p["b"], residual=residual, act=act, colstats=cs)
            y._fd_colstats = cs
            return y
    if want_stats and not pack.has_lora:
        p = pack.pack()
        if arena is not None:
            stats = arena.take(x.shape[0])
            return raw.gemm(x, p["w"], bias=p["b"], residual=residual, rowstats=stats, act=act,
                            rowstats_prezeroed=True), stats
        stats = torch.empty((x.shape[0], 2), device=x.device, dtype=torch.float32)
        return raw.gemm(x, p["w"], bias=p["b"], residual=residual, rowstats=stats, act=act), stats
    y = _linear_fwd_raw(x, pack, residual, act=act, out_fp32=out_fp32)[0]
    return y if want_stats is None else (y, None)


def ln_foldable(x, stats, pack: LinearPack):
    return stats is not None and not pack.has_lora and not _grad_on(x)


def linear_ln(x, stats, norm, pack: LinearPack):
    """LayerNorm(x) W^T + b with the LayerNorm folded into the GEMM epilogue (no LayerNorm pass over memory).
    `stats` are the row sums produced by the GEMM that wrote x."""
    p = pack.pack_ln(norm)
    return raw.gemm(x, p["w"], bias=p["b"], geglu=pack.geglu, ln=(stats, p["colsum"], x.shape[1], norm.eps))


class _GegluFn(torch.autograd.Function):
    """out = value * gelu(gate) with [value|gate] = x W^T + b; backward recomputes the pre-activation."""

    @staticmethod
    def forward(ctx, x, pack):
        p = pack.pack()
        ctx.pack = pack
        ctx.save_for_backward(x)
        return raw.gemm(x, p["w"], bias=p["b"], geglu=True)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        p = ctx.pack.pack()
        acc = raw.gemm(x, p["w"], bias=p["b"])                 # recompute (interleaved layout)
        dacc = raw.geglu_bwd(acc, dout.contiguous())
        return raw.gemm(dacc, ctx.pack.pack_t()), None


def geglu(x, pack: LinearPack):
    if _grad_on(x):
        return _GegluFn.apply(x, pack)
    p = pack.pack()
    return raw.gemm(x, p["w"], bias=p["b"], geglu=True)


# ------------------------------------------------------------------------------------------------
# DiT / MMDiT blocks: AdaLN modulation and AdaLN-Zero gates (gradients flow to scale / shift / gate because the
# reference's SD3 LoRA targets include every AdaLN linear, examples/train_flash_sd3.py:104-117)
# ------------------------------------------------------------------------------------------------
class _ModulateFn(torch.autograd.Function):
    """y = LN(x) * (1 + scale[b]) + shift[b]; scale / shift fp32 [B, C] (views of the AdaLN projection)."""

    @staticmethod
    def forward(ctx, x, scale, shift, rows_per_batch, eps):
        ctx.save_for_backward(x, scale)
        ctx.meta = (rows_per_batch, eps)
        return raw.layernorm_modulate(x, scale, shift, rows_per_batch, eps)

    @staticmethod
    def backward(ctx, dy):
        x, scale = ctx.saved_tensors
        rows_per_batch, eps = ctx.meta
        dx, dscale, dshift = raw.layernorm_modulate_bwd(x, dy.contiguous(), scale, rows_per_batch, eps)
        return dx, dscale, dshift, None, None


def modulate(x, scale, shift, rows_per_batch, eps=1e-6):
    if _grad_on(x, scale, shift):
        return _ModulateFn.apply(x, scale, shift, rows_per_batch, eps)
    return raw.layernorm_modulate(x, scale, shift, rows_per_batch, eps)


class _GateResidualFn(torch.autograd.Function):
    """out = res + gate[b] * h; gate fp32 [B, C]."""

    @staticmethod
    def forward(ctx, h, gate, res, rows_per_batch):
        ctx.save_for_backward(h, gate)
        ctx.rows_per_batch = rows_per_batch
        return raw.gate_residual(h, gate, res, rows_per_batch)

    @staticmethod
    def backward(ctx, dout):
        h, gate = ctx.saved_tensors
        dout = dout.contiguous()
        dh, dgate = raw.gate_bwd(dout, h, gate, ctx.rows_per_batch)
        return dh, dgate, dout, None


def gated_linear(x, pack: LinearPack, gate, res, rows_per_batch, act=0):
    """res + gate[b] * act(x W^T + b (+LoRA)).  Without gradients the gate and the residual ride in the GEMM
    epilogue; with gradients the branch output is kept for the gate gradient."""
    lora_params = pack.lora_params() if pack.has_lora else []
    if _grad_on(x, gate, res, *lora_params):
        h = linear(x, pack, act=act)
        return _GateResidualFn.apply(h, gate, res, rows_per_batch)
    p = pack.pack()
    a2 = b2 = None
    if pack.has_lora:
        lp = pack.pack_lora()
        a2, b2 = raw.gemm(x, lp["a"]), lp["b"]
    return raw.gemm(x, p["w"], a2=a2, b2=b2, bias=p["b"], act=act, residual=res, rowscale=gate,
                    rows_per_group_scale=rows_per_batch)


class _UnpatchifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, NB, h, w, p, Cout, Ckeep):
        ctx.meta = (h, w, p, Cout)
        return raw.unpatchify(x, NB, h, w, p, Cout, Ckeep)

    @staticmethod
    def backward(ctx, dy):
        h, w, p, Cout = ctx.meta
        return raw.patchify(dy.contiguous().float(), h, w, p, Cout).float(), None, None, None, None, None, None


def unpatchify(x, NB, h, w, p, Cout, Ckeep):
    if _grad_on(x):
        return _UnpatchifyFn.apply(x, NB, h, w, p, Cout, Ckeep)
    return raw.unpatchify(x, NB, h, w, p, Cout, Ckeep)


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
class _AttnSelfFn(torch.autograd.Function):
    """Self-attention on the fused projection output qkv [B, N, 3*H*d] (q | k | v)."""

    @staticmethod
    def forward(ctx, qkv, H, head_dim, scale):
        inner = H * head_dim
        q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
        o, lse = raw.attention_fwd(q, k, v, H, scale=scale, need_lse=True, head_dim=head_dim)
        ctx.save_for_backward(qkv, o, lse)
        ctx.meta = (H, head_dim, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        H, head_dim, scale = ctx.meta
        inner = H * head_dim
        q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
        dqkv = torch.empty_like(qkv)
        raw.attention_bwd(q, k, v, o, lse, do.contiguous(), H, scale=scale, dq=dqkv[..., :inner],
                          dk=dqkv[..., inner:2 * inner], dv=dqkv[..., 2 * inner:], head_dim=head_dim)
        return dqkv, None, None, None


class _AttnCrossFn(torch.autograd.Function):
    """Cross-attention: q [B, Nq, H*d], fused kv [B, Nkv, 2*H*d] (k | v), optional key-padding mask kv_len [B]."""

    @staticmethod
    def forward(ctx, q, kv, H, head_dim, scale, kv_len):
        inner = H * head_dim
        o, lse = raw.attention_fwd(q, kv[..., :inner], kv[..., inner:], H, scale=scale, need_lse=True,
                                   head_dim=head_dim, kv_len=kv_len)
        ctx.save_for_backward(q, kv, o, lse)
        ctx.meta = (H, head_dim, scale, kv_len)
        return o

    @staticmethod
    def backward(ctx, do):
        q, kv, o, lse = ctx.saved_tensors
        H, head_dim, scale, kv_len = ctx.meta
        inner = H * head_dim
        dkv = torch.empty_like(kv)
        dq, _, _ = raw.attention_bwd(q, kv[..., :inner], kv[..., inner:], o, lse, do.contiguous(), H, scale=scale,
                                     dk=dkv[..., :inner], dv=dkv[..., inner:], head_dim=head_dim, kv_len=kv_len)
        return dq, dkv, None, None, None, None


def attention_self(qkv, H, head_dim=64, scale=None):
    """qkv [B, N, 3*H*d] -> [B, N, H*d]"""
    if _grad_on(qkv):
        return _AttnSelfFn.apply(qkv, H, head_dim, scale)
    inner = H * head_dim
    return raw.attention_fwd(qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:], H, scale=scale,
                             head_dim=head_dim)


def attention_cross(q, kv, H, head_dim=64, scale=None, kv_len=None):
    """q [B, Nq, H*d], kv [B, Nkv, 2*H*d] -> [B, Nq, H*d]; kv_len [B] int32 masks padded keys."""
    if _grad_on(q, kv):
        return _AttnCrossFn.apply(q, kv, H, head_dim, scale, kv_len)
    inner = H * head_dim
    return raw.attention_fwd(q, kv[..., :inner], kv[..., inner:], H, scale=scale, head_dim=head_dim, kv_len=kv_len)


class _AttnBigHeadFn(torch.autograd.Function):
    """Single-head attention with a head dim beyond the fused kernels' TMEM budget (the VAE mid block: 512 channels):
    per image S = Q K^T and O = P V on fd_gemm around the row-softmax kernel; backward = the five matching GEMMs."""

    @staticmethod
    def forward(ctx, q, k, v, B, scale):
        rows, C = q.shape
        N = rows // B
        outs, probs = [], []
        keep = any(ctx.needs_input_grad[:3])
        for b in range(B):
            sl = slice(b * N, (b + 1) * N)
            P = raw.softmax_rows(raw.gemm(q[sl], k[sl]), scale)                # [N, N]
            outs.append(raw.gemm(P, raw.transpose(v[sl].contiguous())))        # [N, C]
            if keep:
                probs.append(P)
        ctx.meta = (B, N, scale)
        ctx.save_for_backward(q, k, v, *probs)
        return torch.cat(outs, dim=0) if B > 1 else outs[0]

    @staticmethod
    def backward(ctx, do):
        q, k, v, *probs = ctx.saved_tensors
        B, N, scale = ctx.meta
        do = do.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        for b in range(B):
            sl = slice(b * N, (b + 1) * N)
            P, dob = probs[b], do[sl]
            dP = raw.gemm(dob, v[sl])                                           # dO V^T   [N, N]
            raw.gemm(raw.transpose(P), raw.transpose(dob), out=dv[sl])          # P^T dO   [N, C]
            dS = raw.softmax_rows_bwd(P, dP, scale)
            raw.gemm(dS, raw.transpose(k[sl].contiguous()), out=dq[sl])         # dS K
            raw.gemm(raw.transpose(dS), raw.transpose(q[sl].contiguous()), out=dk[sl])   # dS^T Q
        return dq, dk, dv, None, None


def attention_bighead(q, k, v, B, scale):
    """q, k, v [B*N, C] bf16 (views with unit column stride) -> [B*N, C]; one head of C channels."""
    return _AttnBigHeadFn.apply(q, k, v, B, scale)


# ------------------------------------------------------------------------------------------------
# layout ops
# ------------------------------------------------------------------------------------------------
class _ConcatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.c1, ctx.c2 = a.shape[1], b.shape[1]
        return raw.concat_channels(a, b)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        return raw.slice_channels(dy, 0, ctx.c1), raw.slice_channels(dy, ctx.c1, ctx.c2)


def concat(a, b):
    if _grad_on(a, b):
        return _ConcatFn.apply(a, b)
    y = raw.concat_channels(a, b)
    ca, cb = colstats_of(a), colstats_of(b)
    if ca is not None and cb is not None:
        y._fd_colstats = torch.cat([ca, cb], dim=1)
    return y


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, geom):
        ctx.geom = geom
        NB, H, W = geom
        return raw.upsample2x(x, NB, H, W, x.shape[1])

    @staticmethod
    def backward(ctx, dy):
        NB, H, W = ctx.geom
        return raw.upsample2x_bwd(dy.contiguous(), NB, H, W, dy.shape[1]), None


def upsample2x(x, geom):
    if _grad_on(x):
        return _UpsampleFn.apply(x, geom)
    NB, H, W = geom
    return raw.upsample2x(x, NB, H, W, x.shape[1])


class _ToNHWCFn(torch.autograd.Function):
    """NCHW fp32 -> NHWC bf16 (channel-padded); backward returns NCHW fp32."""

    @staticmethod
    def forward(ctx, x, cpad):
        ctx.shape = x.shape
        NB, C, H, W = x.shape
        return raw.nchw_to_nhwc(x, cpad).view(NB * H * W, cpad)

    @staticmethod
    def backward(ctx, dy):
        NB, C, H, W = ctx.shape
        return raw.nhwc_to_nchw(dy.contiguous(), NB, C, H, W), None


def to_nhwc(x, cpad):
    if _grad_on(x):
        return _ToNHWCFn.apply(x.float(), cpad)
    NB, C, H, W = x.shape
    return raw.nchw_to_nhwc(x.float(), cpad).view(NB * H * W, cpad)


class _ToNCHWFn(torch.autograd.Function):
    """NHWC rows [M, ld] (fp32 or bf16, first C valid) -> NCHW fp32."""

    @staticmethod
    def forward(ctx, x, geom, C):
        NB, H, W = geom
        ctx.meta = (geom, C, x.shape[1], x.dtype)
        return raw.nhwc_to_nchw(x, NB, C, H, W)

    @staticmethod
    def backward(ctx, dy):
        (NB, H, W), C, ld, dtype = ctx.meta
        g = raw.nchw_to_nhwc(dy.contiguous().float(), ld).view(NB * H * W, ld)
        if dtype == torch.float32:
            g = g.float()
        return g, None, None


def to_nchw(x, geom, C):
    if _grad_on(x):
        return _ToNCHWFn.apply(x, geom, C)
    NB, H, W = geom
    return raw.nhwc_to_nchw(x, NB, C, H, W)
