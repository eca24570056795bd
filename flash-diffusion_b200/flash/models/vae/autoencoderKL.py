"""B200-native `AutoencoderKLDiffusers` (drop-in for reference src/flash/models/vae/autoencoderKL.py:9-128).

The reference wraps `diffusers.models.AutoencoderKL.from_pretrained(version, subfolder, revision)`; here `AutoencoderKL`
rebuilds the same module tree with the same state-dict keys (`encoder.down_blocks.0.resnets.0.norm1.weight`,
`decoder.mid_block.attentions.0.to_q.weight`, `quant_conv.weight`, ...) as parameter containers, and `encode` /
`decode` walk it on the SAME hand-written sm_100a kernels as the denoisers (BASELINE north_star: "the frozen teacher
forward and VAE encode run on the same kernels"): implicit-GEMM 3x3 convs (the encoder's stride-2 convs with
diffusers' asymmetric right/bottom padding as 9 taps over the space-to-depth image), GroupNorm(+SiLU), and the
single-head 512-channel mid-block attention as two tcgen05 GEMMs around a row softmax (flash.b200.ops).  The decoder is
differentiable with respect to its input (the LPIPS distillation loss back-propagates through it, reference
flash_diffusion_model.py:383-397); the encoder runs frozen, without a graph.  No CPU / eager fallback.

There is no network: `from_pretrained` builds the architecture of the checkpoints the example scripts name with random
weights (load real ones with `load_state_dict`, diffusers keys).  Math restated in oracle/vae.py.
"""
import math

import torch
import torch.nn as nn

from ...b200 import ops
from ...b200.ops import ConvPack, LinearPack
from ..base.base_model import BaseModel
from ..unets.unet import _Container
from .autoencoderKL_config import AutoencoderKLDiffusersConfig

# architecture + scaling factor of the checkpoints named by the example scripts / the reference's own test
# (examples/train_flash_sd.py:174-178, train_flash_sdxl.py, tests/test_vaes/test_autoencoderKL.py:13-24)
_KNOWN = {
    ("runwayml/stable-diffusion-v1-5", "vae"): dict(scaling_factor=0.18215),
    ("stabilityai/sdxl-vae", ""): dict(scaling_factor=0.13025),
    ("stabilityai/stable-diffusion-xl-base-1.0", "vae"): dict(scaling_factor=0.13025),
    ("PixArt-alpha/PixArt-XL-2-1024-MS", "vae"): dict(scaling_factor=0.13025),
    # SD3: 16 latent channels, no quant / post-quant 1x1 convs (examples/train_flash_sd3.py:88-96; the reference
    # wrapper applies `scaling_factor` only and never the checkpoint's shift_factor, vae/autoencoderKL.py:59,78)
    ("stabilityai/stable-diffusion-3-medium", "vae"): dict(scaling_factor=1.5305, latent_channels=16,
                                                           use_quant_conv=False, use_post_quant_conv=False),
}


class VaeResnetBlock(_Container):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None


class VaeAttention(_Container):
    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(ch, ch), nn.Linear(ch, ch), nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])


class VaeMidBlock(_Container):
    def __init__(self, ch, groups):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnetBlock(ch, ch, groups), VaeResnetBlock(ch, ch, groups)])
        self.attentions = nn.ModuleList([VaeAttention(ch, groups)])


class _Sampler(_Container):
    def __init__(self, ch, stride):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=stride, padding=0 if stride == 2 else 1)


class VaeBlock(_Container):
    def __init__(self, cin, cout, layers, groups, down=False, up=False):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnetBlock(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([_Sampler(cout, 2)]) if down else None
        self.upsamplers = nn.ModuleList([_Sampler(cout, 1)]) if up else None


class VaeEncoder(_Container):
    def __init__(self, in_channels, latent_channels, boc, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, co in enumerate(boc):
            self.down_blocks.append(VaeBlock(c, co, layers, groups, down=i != len(boc) - 1))
            c = co
        self.mid_block = VaeMidBlock(c, groups)
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent_channels, 3, padding=1)


class VaeDecoder(_Container):
    def __init__(self, out_channels, latent_channels, boc, layers, groups):
        super().__init__()
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = VaeMidBlock(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(VaeBlock(c, co, layers + 1, groups, up=i != len(rev) - 1))
            c = co
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, out_channels, 3, padding=1)


class _FusedConv(nn.Module):
    """`quant_conv(conv_out(x))`: a 1x1 convolution applied to a 3x3 convolution is one 3x3 convolution with
    W'[o] = sum_m Wq[o, m] W[m], b' = Wq b + bq (exact).  Parameter-less view over the two modules."""

    def __init__(self, conv3, conv1):
        super().__init__()
        self.__dict__["conv3"], self.__dict__["conv1"] = conv3, conv1

    @property
    def weight(self):
        wq = self.conv1.weight.detach().float().reshape(self.conv1.weight.shape[0], -1)
        return torch.einsum("om,mikl->oikl", wq, self.conv3.weight.detach().float())

    @property
    def bias(self):
        wq = self.conv1.weight.detach().float().reshape(self.conv1.weight.shape[0], -1)
        return wq @ self.conv3.bias.detach().float() + self.conv1.bias.detach().float()


class AutoencoderKL(nn.Module):
    """diffusers `AutoencoderKL` (constructor kwargs of the SD / SDXL VAE config) on the B200 kernels."""

    def __init__(self, in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215, latents_mean=None, latents_std=None,
                 use_quant_conv=True, use_post_quant_conv=True, **unused):
        super().__init__()
        boc = list(block_out_channels)
        self.encoder = VaeEncoder(in_channels, latent_channels, boc, layers_per_block, norm_num_groups)
        self.decoder = VaeDecoder(out_channels, latent_channels, boc, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1) if use_quant_conv else None
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1) if use_post_quant_conv else None
        from types import SimpleNamespace
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                                      block_out_channels=tuple(boc), layers_per_block=layers_per_block,
                                      norm_num_groups=norm_num_groups, scaling_factor=scaling_factor,
                                      latents_mean=latents_mean, latents_std=latents_std)
        self.__dict__["_packs"] = {}

    @classmethod
    def from_pretrained(cls, version, subfolder="", revision="main", **kw):
        key = (version, subfolder or "")
        if key not in _KNOWN:
            raise ValueError(f"unknown VAE checkpoint {key}: offline build, known: {sorted(_KNOWN)}")
        return cls(**dict(_KNOWN[key], **kw))

    def _pack(self, key, make):
        packs = self.__dict__.setdefault("_packs", {})
        if key not in packs:
            packs[key] = make()
        return packs[key]

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_packs" else copy.deepcopy(v, memo)
        for m in new.modules():
            m.__dict__.pop("_fd_cache", None)
        return new

    # ------------------------------------------------------------------------------------ engine pieces
    def _resnet(self, r, x, geom):
        c1 = self._pack(("c1", id(r)), lambda: ConvPack(r.conv1))
        c2 = self._pack(("c2", id(r)), lambda: ConvPack(r.conv2, r.conv_shortcut))
        h = ops.group_norm(x, geom, r.norm1, silu=True)
        h = ops.conv3x3(h, geom, c1)
        h = ops.group_norm(h, geom, r.norm2, silu=True)
        if r.conv_shortcut is not None:
            return ops.conv3x3(h, geom, c2, x2=x)
        return ops.conv3x3(h, geom, c2, residual=x)

    def _attention(self, a, x, geom):
        C = x.shape[1]
        h = ops.group_norm(x, geom, a.group_norm, silu=False)
        qkv = ops.linear(h, self._pack(("qkv", id(a)), lambda: LinearPack([a.to_q, a.to_k, a.to_v])))
        o = ops.attention_bighead(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], geom[0], C ** -0.5)
        return ops.linear(o, self._pack(("o", id(a)), lambda: LinearPack(a.to_out[0])), residual=x)

    def _mid(self, mb, x, geom):
        x = self._resnet(mb.resnets[0], x, geom)
        x = self._attention(mb.attentions[0], x, geom)
        return self._resnet(mb.resnets[1], x, geom)

    @staticmethod
    def _check(x):
        if not x.is_cuda:
            raise RuntimeError("AutoencoderKL runs only on CUDA (B200) tensors: there is no CPU fallback")

    # ------------------------------------------------------------------------------------ encoder (frozen, no graph)
    @torch.no_grad()
    def moments(self, x):
        """(mean, logvar) [B, latent, H/8, W/8] fp32 of the posterior; logvar clamped to [-30, 20] as upstream."""
        self._check(x)
        enc = self.encoder
        NB, C, H, W = x.shape
        f = 2 ** (len(enc.down_blocks) - 1)
        if H % f or W % f:
            raise ValueError(f"image size {(H, W)} must be a multiple of {f}")
        conv_in = self._pack("e_in", lambda: ConvPack(enc.conv_in))
        geom = (NB, H, W)
        h = ops.conv3x3(ops.to_nhwc(x, conv_in.cin), geom, conv_in)
        for blk in enc.down_blocks:
            for r in blk.resnets:
                h = self._resnet(r, h, geom)
            if blk.downsamplers is not None:
                ds = blk.downsamplers[0]
                h = ops.conv3x3(h, geom, self._pack(("ds", id(ds)), lambda: ConvPack(ds.conv)), stride=2,
                                pad_mode="asym")
                geom = (NB, geom[1] // 2, geom[2] // 2)
        h = self._mid(enc.mid_block, h, geom)
        h = ops.group_norm(h, geom, enc.conv_norm_out, silu=True)
        head = self._pack("e_out", lambda: ConvPack(enc.conv_out if self.quant_conv is None
                                                    else _FusedConv(enc.conv_out, self.quant_conv)))
        m = ops.to_nchw(ops.conv3x3(h, geom, head, out_fp32=True), geom, 2 * self.config.latent_channels)
        mean, logvar = m.chunk(2, dim=1)
        return mean, logvar.clamp(-30.0, 20.0)

    def encode_sample(self, x, noise=None):
        """posterior sample (UPSTREAM `vae.encode(x).latent_dist.sample()`); `noise` makes the draw explicit."""
        mean, logvar = self.moments(x)
        if noise is None:
            noise = torch.randn_like(mean)
        return mean + torch.exp(0.5 * logvar) * noise

    # ------------------------------------------------------------------------------------ decoder (input gradient)
    def decode(self, z):
        """UPSTREAM `vae.decode(z).sample`: [B, latent, h, w] fp32 -> [B, 3, 8h, 8w] fp32."""
        self._check(z)
        dec = self.decoder
        NB, C, H, W = z.shape
        # post_quant_conv: a 4x4 channel mix of the fp32 latent at the NCHW boundary (16 FMAs per pixel)
        if self.post_quant_conv is not None:
            wq = self.post_quant_conv.weight.detach().float().reshape(C, C)
            z = torch.einsum("oi,bihw->bohw", wq, z.float()) + self.post_quant_conv.bias.detach().float().view(1, C, 1, 1)
        conv_in = self._pack("d_in", lambda: ConvPack(dec.conv_in))
        geom = (NB, H, W)
        h = ops.conv3x3(ops.to_nhwc(z, conv_in.cin), geom, conv_in)
        h = self._mid(dec.mid_block, h, geom)
        for blk in dec.up_blocks:
            for r in blk.resnets:
                h = self._resnet(r, h, geom)
            if blk.upsamplers is not None:
                us = blk.upsamplers[0]
                h = ops.upsample2x(h, geom)
                geom = (NB, geom[1] * 2, geom[2] * 2)
                h = ops.conv3x3(h, geom, self._pack(("us", id(us)), lambda: ConvPack(us.conv)))
        h = ops.group_norm(h, geom, dec.conv_norm_out, silu=True)
        conv_out = self._pack("d_out", lambda: ConvPack(dec.conv_out))
        return ops.to_nchw(ops.conv3x3(h, geom, conv_out, out_fp32=True), geom, self.config.out_channels)


class AutoencoderKLDiffusers(BaseModel):
    """Same constructor, attributes and `encode` / `decode` contract as the reference wrapper
    (src/flash/models/vae/autoencoderKL.py:9-128)."""

    def __init__(self, config: AutoencoderKLDiffusersConfig):
        BaseModel.__init__(self, config)
        self.config = config
        self.vae_model = AutoencoderKL.from_pretrained(config.version, subfolder=config.subfolder,
                                                       revision=config.revision)
        self.tiling_size = config.tiling_size
        self.tiling_overlap = config.tiling_overlap
        # reference `_get_properties` (:32-50) measures the factor with a 32x32 probe on the CPU; the kernels are
        # CUDA-only, so it is derived from the architecture (one stride-2 conv per block but the last)
        self.downsampling_factor = 2 ** (len(self.vae_model.config.block_out_channels) - 1)
        self.latent_channels = self.vae_model.config.latent_channels
        self.latents_mean = self.vae_model.config.latents_mean
        self.latents_std = self.vae_model.config.latents_std
        self.has_latents_mean = self.latents_mean is not None
        self.has_latents_std = self.latents_std is not None

    def encode(self, x: torch.Tensor, batch_size: int = 8, noise: torch.Tensor = None):
        """reference :52-62 — posterior sample per chunk of `batch_size` images, times `scaling_factor`."""
        latents = []
        for i in range(0, x.shape[0], batch_size):
            n = None if noise is None else noise[i:i + batch_size]
            latents.append(self.vae_model.encode_sample(x[i:i + batch_size], n))
        return torch.cat(latents, dim=0) * self.vae_model.config.scaling_factor

    def decode(self, z: torch.Tensor):
        """reference :64-128 — un-scale, then decode; latents larger than `tiling_size` are decoded tile by tile with
        `tiling_overlap` and blended (the reference merges the tiles on the CPU and returns a CPU tensor, SURVEY Q12;
        here the result stays on the device)."""
        sf = self.vae_model.config.scaling_factor
        if self.has_latents_mean and self.has_latents_std:
            mean = torch.tensor(self.latents_mean).view(1, self.latent_channels, 1, 1).to(z.device, z.dtype)
            std = torch.tensor(self.latents_std).view(1, self.latent_channels, 1, 1).to(z.device, z.dtype)
            z = z * std / sf + mean
        else:
            z = z / sf
        th, tw = self.tiling_size
        if z.shape[2] <= th and z.shape[3] <= tw:
            return self.vae_model.decode(z)
        return self._decode_tiled(z)

    @staticmethod
    def _tile_window(n, centre, device):
        """Gaussian tile weight of the reference's `Tiler._gaussian_weights` (src/flash/models/utils.py:156-204):
        exp(-(t - centre)^2 / n^2 / (2 var)) / sqrt(2 pi var), var = 0.01 — evaluated in float64 as numpy does."""
        t = torch.arange(n, dtype=torch.float64, device=device)
        var = 0.01
        return torch.exp(-(t - centre) ** 2 / (n * n) / (2 * var)) / math.sqrt(2 * math.pi * var)

    @torch.no_grad()
    def _decode_tiled(self, z):
        """reference :80-124 — `Tiler.get_tiles` (stride = tile - overlap, the overlap only along an axis that is
        actually tiled, trailing tiles may be partial), every tile zero-padded to the tile size, decoded, cropped,
        then `Tiler.merge_tiles("gaussian")`: sum(tile * w) / sum(w) with the gaussian window of the CROPPED tile (its
        column window is centred on (n - 1) / 2, its row window on n / 2, as upstream).  On the device, whole batch at
        once (the reference loops over samples and merges on the CPU)."""
        th, tw = self.tiling_size
        oh = self.tiling_overlap[0] if z.shape[2] > th else 0
        ow = self.tiling_overlap[1] if z.shape[3] > tw else 0
        f = self.downsampling_factor
        B, _, H, W = z.shape
        out = torch.zeros((B, self.vae_model.config.out_channels, H * f, W * f), device=z.device, dtype=torch.float64)
        wsum = torch.zeros((1, 1, H * f, W * f), device=z.device, dtype=torch.float64)
        for i in range(0, H, th - oh):
            for j in range(0, W, tw - ow):
                tile = z[:, :, i:i + th, j:j + tw]
                hh, ww = tile.shape[2], tile.shape[3]
                pad = torch.zeros((B, z.shape[1], th, tw), device=z.device, dtype=z.dtype)
                pad[:, :, :hh, :ww] = tile                    # decode at the fixed tile size (reference :97-104)
                dec = self.vae_model.decode(pad)[:, :, :hh * f, :ww * f]
                wy = self._tile_window(hh * f, hh * f / 2, z.device)
                wx = self._tile_window(ww * f, (ww * f - 1) / 2, z.device)
                w2 = wy[:, None] * wx[None, :]
                out[:, :, i * f:i * f + hh * f, j * f:j * f + ww * f] += dec.double() * w2
                wsum[:, :, i * f:i * f + hh * f, j * f:j * f + ww * f] += w2
        return (out / wsum).float()
