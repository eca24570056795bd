"""ORACLE (test infrastructure — never imported by the product path).

fp32 PyTorch restatement of the arithmetic behind the reference's PixArt-alpha wrapper
`DiffusersTransformer2DWrapper` (reference src/flash/models/transformers/tranformers.py:9-100 and the custom
`AdaLayerNormSingle`, src/flash/models/transformers/utils.py:8-102), i.e. diffusers' `Transformer2DModel` with
`norm_type="ada_norm_single"`, `patch_size=2` as constructed at examples/train_flash_pixart.py:65-86.

PARITY UNPINNED (see oracle/unet.py): the UPSTREAM module math (PatchEmbed + 2D sin-cos position table,
PixArtAlphaTextProjection, BasicTransformerBlock(ada_norm_single), output AdaLN, un-patchify) is restated from the
published diffusers implementation; state-dict keys follow the names the reference pokes
(examples/train_flash_pixart.py:90-172: `adaln_single.timestep_embedder.linear_{1,2}`,
`adaln_single.add_embedding.<i>.linear_{1,2}`).
"""
from typing import Dict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import TimestepEmbedding, timestep_embedding


def sincos_1d(embed_dim, pos):
    omega = np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim, grid_size, base_size, interpolation_scale):
    """diffusers get_2d_sincos_pos_embed (w first in the meshgrid, [emb_h | emb_w] concatenation)."""
    rows, cols = (grid_size, grid_size) if np.isscalar(grid_size) else grid_size
    ys = np.arange(rows, dtype=np.float32) / (rows / base_size) / interpolation_scale
    xs = np.arange(cols, dtype=np.float32) / (cols / base_size) / interpolation_scale
    grid = np.stack(np.meshgrid(xs, ys), axis=0).reshape(2, 1, rows, cols)
    emb_h = sincos_1d(embed_dim // 2, grid[0])
    emb_w = sincos_1d(embed_dim // 2, grid[1])
    return np.concatenate([emb_h, emb_w], axis=1)


class PatchEmbed(nn.Module):
    def __init__(self, sample_size, patch_size, in_channels, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, patch_size, stride=patch_size)
        grid = sample_size // patch_size
        self.patch, self.grid, self.interp = patch_size, grid, max(sample_size // 64, 1)
        pe = sincos_2d(embed_dim, grid, base_size=grid, interpolation_scale=self.interp)
        self.register_buffer("pos_embed", torch.from_numpy(pe).float()[None], persistent=False)

    def forward(self, x):
        rows, cols = x.shape[-2] // self.patch, x.shape[-1] // self.patch
        x = self.proj(x).flatten(2).transpose(1, 2)
        if (rows, cols) != (self.grid, self.grid):      # diffusers PatchEmbed.forward: table recomputed for (h, w)
            pe = sincos_2d(self.pos_embed.shape[-1], (rows, cols), base_size=self.grid, interpolation_scale=self.interp)
            return x + torch.from_numpy(pe).float()[None].to(x)
        return x + self.pos_embed.to(x.dtype)


class AdaLayerNormSingle(nn.Module):
    """reference src/flash/models/transformers/utils.py:8-102"""

    def __init__(self, time_embed_dim, timesteps_embedding_num_channels=256, projection_class_embeddings_input_dim=None,
                 use_concat_conditioning=False, num_vector_conditionings=None):
        super().__init__()
        self.nch = timesteps_embedding_num_channels
        self.timestep_embedder = TimestepEmbedding(timesteps_embedding_num_channels, time_embed_dim)
        self.in_dim = projection_class_embeddings_input_dim
        self.nvec = num_vector_conditionings
        if self.in_dim is not None:
            if not use_concat_conditioning:
                self.add_embedding = TimestepEmbedding(self.in_dim, time_embed_dim)
            else:
                self.add_embedding = nn.ModuleList(
                    [TimestepEmbedding(self.in_dim, time_embed_dim // num_vector_conditionings)
                     for _ in range(num_vector_conditionings)])
        self.linear = nn.Linear(time_embed_dim, 6 * time_embed_dim)

    def forward(self, timestep, vector):
        emb = self.timestep_embedder(timestep_embedding(timestep.reshape(-1), self.nch))
        if self.in_dim is not None:
            if isinstance(self.add_embedding, nn.ModuleList):
                parts = torch.chunk(vector, self.nvec, dim=1)
                emb = emb + torch.cat([m(p) for m, p in zip(self.add_embedding, parts)], dim=1)
            else:
                emb = emb + self.add_embedding(vector)
        return self.linear(F.silu(emb)), emb


class TextProjection(nn.Module):
    """PixArtAlphaTextProjection: Linear -> GELU(tanh) -> Linear"""

    def __init__(self, in_features, hidden):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden)
        self.linear_2 = nn.Linear(hidden, hidden)

    def forward(self, x):
        return self.linear_2(F.gelu(self.linear_1(x), approximate="tanh"))


class Attention(nn.Module):
    def __init__(self, dim, cross_dim, heads, dim_head, bias):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_dim or dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_dim or dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])

    def forward(self, x, context=None, mask=None):
        context = x if context is None else context
        B, N, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(context), self.to_v(context)
        d = q.shape[-1] // self.heads
        q = q.view(B, N, self.heads, d).transpose(1, 2)
        k = k.view(B, -1, self.heads, d).transpose(1, 2)
        v = v.view(B, -1, self.heads, d).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) * d ** -0.5
        if mask is not None:                       # [B, T] 1 = keep; diffusers adds (1 - mask) * -10000
            s = s + ((1 - mask.to(s.dtype)) * -10000.0)[:, None, None, :]
        o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, N, self.heads * d)
        return self.to_out[0](o)


class GELUProj(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class GEGLUProj(nn.Module):
    """diffusers `GEGLU`: value, gate = proj(x).chunk(2, -1); value * gelu(gate)  (exact erf GELU)"""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, 2 * dim_out)

    def forward(self, x):
        v, g = self.proj(x).chunk(2, dim=-1)
        return v * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, geglu=False):
        super().__init__()
        first = GEGLUProj(dim, dim * mult) if geglu else GELUProj(dim, dim * mult)
        self.net = nn.ModuleList([first, nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class AdaBlock(nn.Module):
    """BasicTransformerBlock(norm_type="ada_norm_single")"""

    def __init__(self, dim, heads, dim_head, cross_dim, bias, eps, affine=False, geglu=False,
                 double_self_attention=False):
        super().__init__()
        self.scale_shift_table = nn.Parameter(torch.randn(6, dim) / dim ** 0.5)
        self.eps = eps
        # UPSTREAM: norm1 / norm2 are nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine); attn2 exists when
        # there is a cross_attention_dim or `double_self_attention` (then it attends to the hidden states themselves)
        self.norm1 = nn.LayerNorm(dim, eps=eps, elementwise_affine=True) if affine else None
        self.attn1 = Attention(dim, None, heads, dim_head, bias)
        self.has_attn2 = cross_dim is not None or double_self_attention
        self.attn2 = Attention(dim, None if double_self_attention else cross_dim, heads, dim_head, bias) \
            if self.has_attn2 else None
        self.double_self = double_self_attention
        self.norm2 = nn.LayerNorm(dim, eps=eps, elementwise_affine=True) if affine else None
        self.ff = FeedForward(dim, geglu=geglu)

    def _norm(self, norm, h):
        return norm(h) if norm is not None else F.layer_norm(h, (h.shape[-1],), eps=self.eps)

    def forward(self, h, ctx, mask, t6):
        B, _, D = h.shape
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = (self.scale_shift_table[None] + t6.reshape(B, 6, -1)).chunk(6, dim=1)
        n = self._norm(self.norm1, h) * (1 + sc_a) + sh_a
        h = g_a * self.attn1(n) + h
        if self.attn2 is not None:               # ada_norm_single: attn2 sees the un-normalised hidden states
            h = (self.attn2(h) if self.double_self else self.attn2(h, ctx, mask)) + h
        n = self._norm(self.norm2, h) * (1 + sc_m) + sh_m
        return g_m * self.ff(n) + h


class PixArtTransformerOracle(nn.Module):
    """Same constructor kwargs / keys / forward contract as the reference `DiffusersTransformer2DWrapper`."""

    def __init__(self, time_embed_dim=256, timesteps_embedding_num_channels=256, projection_class_embeddings_input_dim=None,
                 use_concat_vector_conditioning=False, num_vector_conditionings=None, sample_size=None, num_layers=1,
                 attention_head_dim=88, in_channels=None, out_channels=None, patch_size=None, attention_bias=False,
                 num_attention_heads=16, cross_attention_dim=None, activation_fn="geglu",
                 norm_type="layer_norm", norm_elementwise_affine=True, norm_eps=1e-5, caption_channels=None,
                 double_self_attention=False, **unused):
        # keyword defaults = diffusers Transformer2DModel's; only the ada_norm_single / patch variants are restated
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        assert norm_type == "ada_norm_single" and activation_fn in ("gelu-approximate", "geglu")
        D = num_attention_heads * attention_head_dim
        self.p, self.out_channels, self.eps = patch_size, out_channels, norm_eps
        self.pos_embed = PatchEmbed(sample_size, patch_size, in_channels, D)
        self.adaln_single = AdaLayerNormSingle(time_embed_dim, timesteps_embedding_num_channels,
                                               projection_class_embeddings_input_dim, use_concat_vector_conditioning,
                                               num_vector_conditionings)
        self.caption_projection = TextProjection(caption_channels, D) if caption_channels is not None else None
        self.transformer_blocks = nn.ModuleList(
            [AdaBlock(D, num_attention_heads, attention_head_dim, cross_attention_dim, attention_bias, norm_eps,
                      affine=bool(norm_elementwise_affine), geglu=activation_fn == "geglu",
                      double_self_attention=double_self_attention)
             for _ in range(num_layers)])
        self.scale_shift_table = nn.Parameter(torch.randn(2, D) / D ** 0.5)
        self.proj_out = nn.Linear(D, patch_size * patch_size * out_channels)

    def forward(self, sample, timestep, conditioning: Dict[str, Dict[str, torch.Tensor]], *args, **kwargs):
        assert isinstance(conditioning, dict), "conditionings must be a dictionary"
        cond = conditioning["cond"]
        vector, ctx, concat, mask = cond.get("vector"), cond.get("crossattn"), cond.get("concat"), cond.get("attention_mask")
        c_in = sample.shape[1]
        if concat is not None:
            sample = torch.cat([sample, concat], dim=1)
        B, _, H, W = sample.shape
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        timestep = timestep.reshape(-1).float().expand(B) if timestep.numel() == 1 else timestep.float()
        h = self.pos_embed(sample)
        t6, emb = self.adaln_single(timestep, vector)
        if self.caption_projection is not None:
            ctx = self.caption_projection(ctx)
        for blk in self.transformer_blocks:
            h = blk(h, ctx, mask, t6)
        shift, scale = (self.scale_shift_table[None] + emb[:, None]).chunk(2, dim=1)
        h = F.layer_norm(h, (h.shape[-1],), eps=1e-6) * (1 + scale) + shift     # UPSTREAM norm_out: eps 1e-6, no affine
        h = self.proj_out(h)
        hh, ww, p, c = H // self.p, W // self.p, self.p, self.out_channels
        h = h.reshape(B, hh, ww, p, p, c)
        h = torch.einsum("nhwpqc->nchpwq", h).reshape(B, c, hh * p, ww * p)
        return h[:, :c_in]                       # reference wrapper: `.sample[:, :sample_channels]`

    def freeze(self):
        self.eval()
        for p_ in self.parameters():
            p_.requires_grad = False


PIXART_KWARGS = dict(   # examples/train_flash_pixart.py:65-86
    sample_size=128, num_layers=28, attention_head_dim=72, in_channels=4, out_channels=8, patch_size=2,
    attention_bias=True, num_attention_heads=16, cross_attention_dim=1152, activation_fn="gelu-approximate",
    num_embeds_ada_norm=1000, norm_type="ada_norm_single", norm_elementwise_affine=False, norm_eps=1e-6,
    caption_channels=4096, projection_class_embeddings_input_dim=256, time_embed_dim=1152,
    timesteps_embedding_num_channels=256, use_concat_vector_conditioning=True, num_vector_conditionings=3)
