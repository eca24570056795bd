"""TEST INFRASTRUCTURE ONLY (the checker, never the product path): fp32 PyTorch restatement of diffusers' UNCONDITIONAL
`UNet2DModel` as the reference wraps it (`DiffusersUNet2DWrapper`, src/flash/models/unets/unet.py:7-52).

Parity unpinned upstream: diffusers is not installable here and the reference holds no vectors for this model (its only
test checks the output shape, tests/test_unet/test_unets_wrappers.py:29-42).  Restated from the published diffusers
0.27 sources (`models/unets/unet_2d.py`, `unet_2d_blocks.py: DownBlock2D / AttnDownBlock2D / UNetMidBlock2D /
AttnUpBlock2D / UpBlock2D`, `attention_processor.Attention` in its `_from_deprecated_attn_block` configuration):

  emb    = time_embedding(sinusoid(t, block_out_channels[0])) [+ class_embedding[labels]   (nn.Embedding)]
  block  = ResnetBlock2D(x, emb) [-> x + to_out(softmax(q k^T / sqrt(d)) v),  q / k / v = Linear(GroupNorm(x)), heads of
           `attention_head_dim` channels, biases on] ; Downsample2D = 3x3 stride-2 conv, Upsample2D = nearest x2 + 3x3 conv
  out    = conv_out(SiLU(GroupNorm(x)))
State-dict keys follow diffusers (`down_blocks.1.attentions.0.group_norm.weight`, `...to_q.bias`, `class_embedding.weight`).
"""
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import Downsample2D, ResnetBlock2D, TimestepEmbedding, Upsample2D, timestep_embedding

DOWN_TYPES = ("DownBlock2D", "AttnDownBlock2D")
UP_TYPES = ("UpBlock2D", "AttnUpBlock2D")


class AttentionBlock(nn.Module):
    """diffusers `Attention(channels, heads=channels // head_dim, dim_head=head_dim, norm_num_groups=groups,
    residual_connection=True, bias=True, rescale_output_factor=1)` on a [B, C, H, W] map."""

    def __init__(self, channels, head_dim, groups, eps):
        super().__init__()
        self.heads = channels // head_dim if head_dim is not None else 1
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).reshape(B, C, H * W).transpose(1, 2)
        d = C // self.heads
        split = lambda t: t.reshape(B, H * W, self.heads, d).transpose(1, 2)
        q, k, v = split(self.to_q(h)), split(self.to_k(h)), split(self.to_v(h))
        p = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, dim=-1)
        o = (p @ v).transpose(1, 2).reshape(B, H * W, C)
        return self.to_out[0](o).transpose(1, 2).reshape(B, C, H, W) + x


class DownBlock(nn.Module):
    def __init__(self, in_ch, out_ch, temb_ch, num_layers, add_downsample, groups, eps, head_dim, attn):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups, eps) for i in range(num_layers)])
        self.attentions = nn.ModuleList([AttentionBlock(out_ch, head_dim, groups, eps) for _ in range(num_layers)]) \
            if attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, x, temb):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb_ch, groups, eps, head_dim, add_attention):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([AttentionBlock(ch, head_dim, groups, eps) if add_attention else None])

    def forward(self, x, temb):
        x = self.resnets[0](x, temb)
        if self.attentions[0] is not None:
            x = self.attentions[0](x)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, in_ch, out_ch, prev_ch, temb_ch, num_layers, add_upsample, groups, eps, head_dim, attn):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_ch if i == num_layers - 1 else out_ch
            rin = prev_ch if i == 0 else out_ch
            resnets.append(ResnetBlock2D(rin + skip, out_ch, temb_ch, groups, eps))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList([AttentionBlock(out_ch, head_dim, groups, eps) for _ in range(num_layers)]) \
            if attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, x, skips: List[torch.Tensor], temb):
        for i, res in enumerate(self.resnets):
            x = res(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DOracle(nn.Module):
    """Constructor keywords and defaults of diffusers `UNet2DModel`; forward contract of the reference wrapper."""

    def __init__(self, sample_size=None, in_channels=3, out_channels=3, center_input_sample=False,
                 time_embedding_type="positional", freq_shift=0, flip_sin_to_cos=True,
                 down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
                 up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels=(224, 448, 672, 896), layers_per_block=2, mid_block_scale_factor=1,
                 downsample_padding=1, downsample_type="conv", upsample_type="conv", dropout=0.0, act_fn="silu",
                 attention_head_dim: Optional[int] = 8, norm_num_groups=32, attn_norm_num_groups=None, norm_eps=1e-5,
                 resnet_time_scale_shift="default", add_attention=True, class_embed_type=None, num_class_embeds=None,
                 num_train_timesteps=None, **unused):
        super().__init__()
        assert time_embedding_type == "positional" and act_fn == "silu" and resnet_time_scale_shift == "default"
        assert downsample_type == "conv" and upsample_type == "conv" and class_embed_type is None
        boc = list(block_out_channels)
        n = len(boc)
        temb_ch = boc[0] * 4
        self.time_dim, self.flip, self.shift = boc[0], flip_sin_to_cos, freq_shift
        self.center_input_sample = center_input_sample
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.class_embedding = nn.Embedding(num_class_embeds, temb_ch) if num_class_embeds is not None else None
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            assert t in DOWN_TYPES, t
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownBlock(in_ch, out_ch, temb_ch, layers_per_block, i != n - 1, norm_num_groups,
                                              norm_eps, attention_head_dim if attention_head_dim is not None else out_ch,
                                              t == "AttnDownBlock2D"))
        self.mid_block = MidBlock(boc[-1], temb_ch, norm_num_groups, norm_eps,
                                  attention_head_dim if attention_head_dim is not None else boc[-1], add_attention)
        self.up_blocks = nn.ModuleList()
        rboc = boc[::-1]
        out_ch = rboc[0]
        for i, t in enumerate(up_block_types):
            assert t in UP_TYPES, t
            prev_ch, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, n - 1)]
            self.up_blocks.append(UpBlock(in_ch, out_ch, prev_ch, temb_ch, layers_per_block + 1, i != n - 1,
                                          norm_num_groups, norm_eps,
                                          attention_head_dim if attention_head_dim is not None else out_ch,
                                          t == "AttnUpBlock2D"))
        groups_out = norm_num_groups if norm_num_groups is not None else min(boc[0] // 4, 32)
        self.conv_norm_out = nn.GroupNorm(groups_out, boc[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, sample, timestep, conditioning: Dict[str, Dict[str, torch.Tensor]] = None, *args, **kwargs):
        class_labels = concat = None
        if conditioning is not None:                       # reference wrapper :33-45
            class_labels = conditioning["cond"].get("vector")
            concat = conditioning["cond"].get("concat")
        if concat is not None:
            sample = torch.cat([sample, concat], dim=1)
        B = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        timestep = timestep.reshape(-1).float().to(sample.device)
        timestep = timestep.expand(B) if timestep.numel() == 1 else timestep
        emb = self.time_embedding(timestep_embedding(timestep, self.time_dim, self.flip, self.shift))
        if self.class_embedding is not None:
            if class_labels is None:
                raise ValueError("class_labels should be provided when doing class conditioning")
            emb = emb + self.class_embedding(class_labels.long())
        if self.center_input_sample:
            sample = 2 * sample - 1.0
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb)
            skips += outs
        x = self.mid_block(x, emb)
        for blk in self.up_blocks:
            x = blk(x, skips, emb)
        return self.conv_out(F.silu(self.conv_norm_out(x)))

    def freeze(self):
        self.eval()
        for p in self.parameters():
            p.requires_grad = False
