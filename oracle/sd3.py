"""ORACLE (test infrastructure — never imported by the product path).

fp32 PyTorch restatement of the arithmetic behind the reference's `DiffusersSD3Transformer2DWrapper`
(reference src/flash/models/transformers/tranformers.py:103-163), i.e. diffusers' `SD3Transformer2DModel`
(MMDiT: joint attention over image + text tokens, AdaLN-Zero modulation) with the kwargs of
examples/train_flash_sd3.py:65-77.  PARITY UNPINNED (see oracle/unet.py); pinned structurally by the parameter count
(SD3-medium ~2.03 B) and the diffusers key names.
"""
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .dit import FeedForward, sincos_2d
from .unet import TimestepEmbedding, timestep_embedding


class PatchEmbedSD3(nn.Module):
    def __init__(self, sample_size, patch_size, in_channels, embed_dim, pos_embed_max_size):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, patch_size, stride=patch_size)
        self.max, self.p = pos_embed_max_size, patch_size
        pe = sincos_2d(embed_dim, pos_embed_max_size, base_size=sample_size // patch_size, interpolation_scale=1)
        self.register_buffer("pos_embed", torch.from_numpy(pe).float()[None], persistent=True)

    def cropped(self, h, w):
        top, left = (self.max - h) // 2, (self.max - w) // 2
        pe = self.pos_embed.reshape(1, self.max, self.max, -1)[:, top:top + h, left:left + w]
        return pe.reshape(1, h * w, -1)

    def forward(self, x):
        h, w = x.shape[-2] // self.p, x.shape[-1] // self.p
        x = self.proj(x).flatten(2).transpose(1, 2)
        return x + self.cropped(h, w).to(x.dtype)


class TextProjSilu(nn.Module):
    def __init__(self, in_features, hidden):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden)
        self.linear_2 = nn.Linear(hidden, hidden)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class CombinedTimestepTextProjEmbeddings(nn.Module):
    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, embedding_dim)
        self.text_embedder = TextProjSilu(pooled_projection_dim, embedding_dim)

    def forward(self, timestep, pooled):
        return self.timestep_embedder(timestep_embedding(timestep, 256)) + self.text_embedder(pooled)


class AdaLinear(nn.Module):
    """AdaLayerNormZero / AdaLayerNormContinuous: only `linear` carries parameters (the LayerNorm has no affine)."""

    def __init__(self, dim, chunks):
        super().__init__()
        self.linear = nn.Linear(dim, chunks * dim)

    def forward(self, temb):
        return self.linear(F.silu(temb))


class JointAttention(nn.Module):
    def __init__(self, dim, heads, dim_head, context_pre_only):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        self.add_k_proj, self.add_v_proj, self.add_q_proj = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])
        self.to_add_out = None if context_pre_only else nn.Linear(inner, dim)

    def forward(self, x, c):
        B, N, _ = x.shape
        q = torch.cat([self.to_q(x), self.add_q_proj(c)], dim=1)
        k = torch.cat([self.to_k(x), self.add_k_proj(c)], dim=1)
        v = torch.cat([self.to_v(x), self.add_v_proj(c)], dim=1)
        d = q.shape[-1] // self.heads
        q, k, v = (t.view(B, -1, self.heads, d).transpose(1, 2) for t in (q, k, v))
        o = (torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, dim=-1) @ v).transpose(1, 2).reshape(B, -1, self.heads * d)
        xo = self.to_out[0](o[:, :N])
        co = self.to_add_out(o[:, N:]) if self.to_add_out is not None else None
        return xo, co


class JointBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, context_pre_only):
        super().__init__()
        self.pre_only = context_pre_only
        self.norm1 = AdaLinear(dim, 6)
        self.norm1_context = AdaLinear(dim, 2 if context_pre_only else 6)
        self.attn = JointAttention(dim, heads, dim_head, context_pre_only)
        self.ff = FeedForward(dim)
        self.ff_context = None if context_pre_only else FeedForward(dim)

    @staticmethod
    def _ln(x):
        return F.layer_norm(x, (x.shape[-1],), eps=1e-6)

    def forward(self, x, c, temb):
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = self.norm1(temb).chunk(6, dim=1)
        nx = self._ln(x) * (1 + sc_a[:, None]) + sh_a[:, None]
        if self.pre_only:
            scale, shift = self.norm1_context(temb).chunk(2, dim=1)
            nc = self._ln(c) * (1 + scale[:, None]) + shift[:, None]
        else:
            csh_a, csc_a, cg_a, csh_m, csc_m, cg_m = self.norm1_context(temb).chunk(6, dim=1)
            nc = self._ln(c) * (1 + csc_a[:, None]) + csh_a[:, None]
        ax, ac = self.attn(nx, nc)
        x = x + g_a[:, None] * ax
        x = x + g_m[:, None] * self.ff(self._ln(x) * (1 + sc_m[:, None]) + sh_m[:, None])
        if self.pre_only:
            return x, None
        c = c + cg_a[:, None] * ac
        c = c + cg_m[:, None] * self.ff_context(self._ln(c) * (1 + csc_m[:, None]) + csh_m[:, None])
        return x, c


class SD3TransformerOracle(nn.Module):
    def __init__(self, sample_size=128, patch_size=2, in_channels=16, num_layers=18, attention_head_dim=64,
                 num_attention_heads=18, joint_attention_dim=4096, caption_projection_dim=1152,
                 pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=96, **unused):
        super().__init__()
        D = num_attention_heads * attention_head_dim
        self.p, self.out_channels = patch_size, out_channels
        self.pos_embed = PatchEmbedSD3(sample_size, patch_size, in_channels, D, pos_embed_max_size)
        self.time_text_embed = CombinedTimestepTextProjEmbeddings(D, pooled_projection_dim)
        self.context_embedder = nn.Linear(joint_attention_dim, caption_projection_dim)
        self.transformer_blocks = nn.ModuleList(
            [JointBlock(D, num_attention_heads, attention_head_dim, context_pre_only=(i == num_layers - 1))
             for i in range(num_layers)])
        self.norm_out = AdaLinear(D, 2)
        self.proj_out = nn.Linear(D, patch_size * patch_size * out_channels)

    def forward(self, sample, timestep, conditioning: Dict[str, Dict[str, torch.Tensor]], *args, **kwargs):
        assert isinstance(conditioning, dict), "conditionings must be a dictionary"
        cond = conditioning["cond"]
        pooled, ctx, concat = cond.get("vector"), cond.get("crossattn"), cond.get("concat")
        c_in = sample.shape[1]
        if concat is not None:
            sample = torch.cat([sample, concat], dim=1)
        B, _, H, W = sample.shape
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        timestep = timestep.reshape(-1).float()
        if timestep.numel() == 1:
            timestep = timestep.expand(B)
        x = self.pos_embed(sample)
        temb = self.time_text_embed(timestep, pooled)
        c = self.context_embedder(ctx)
        for blk in self.transformer_blocks:
            x, c = blk(x, c, temb)
        scale, shift = self.norm_out(temb).chunk(2, dim=1)
        x = F.layer_norm(x, (x.shape[-1],), eps=1e-6) * (1 + scale[:, None]) + shift[:, None]
        x = self.proj_out(x)
        hh, ww, p, co = H // self.p, W // self.p, self.p, self.out_channels
        x = torch.einsum("nhwpqc->nchpwq", x.reshape(B, hh, ww, p, p, co)).reshape(B, co, hh * p, ww * p)
        return x[:, :c_in]

    def freeze(self):
        self.eval()
        for p_ in self.parameters():
            p_.requires_grad = False


SD3_KWARGS = dict(sample_size=128, patch_size=2, in_channels=16, num_layers=24, attention_head_dim=64,
                  num_attention_heads=24, joint_attention_dim=4096, caption_projection_dim=1536,
                  pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=192)   # examples/train_flash_sd3.py:65-77
