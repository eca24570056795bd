"""B200-native `DiffusersUNet2DWrapper` (drop-in for reference src/flash/models/unets/unet.py:7-52, which subclasses
diffusers' UNCONDITIONAL `UNet2DModel`): same constructor keywords / defaults, state-dict keys
(`down_blocks.1.attentions.0.group_norm.weight`, `...to_q.bias`, `class_embedding.weight`, ...) and `forward(sample,
timestep, conditioning=None)` contract (`cond["vector"]` = integer class labels, `cond["concat"]` channel-concatenated).

It runs on the SAME engine as the conditional wrapper (unet.py): ResnetBlock2D = GroupNorm+SiLU -> implicit-GEMM conv
(+time-embedding row) -> GroupNorm+SiLU -> conv (+shortcut K segment / residual epilogue); the spatial self-attention
blocks of `AttnDownBlock2D` / `UNetMidBlock2D` / `AttnUpBlock2D` are GroupNorm -> fused q|k|v GEMM (biases on, heads of
`attention_head_dim` channels zero-padded to the attention kernel's 16-channel granularity) -> FlashAttention kernel ->
out-projection with the residual in the epilogue.  No CPU / eager fallback.  Math restated in oracle/unet2d.py."""
from typing import Dict, Optional, Union

import torch
import torch.nn as nn

from ...b200 import ops, raw
from ...b200.ops import LinearPack
from .unet import (DiffusersUNet2DCondWrapper, Downsample2D, ResnetBlock2D, TimestepEmbedding, Upsample2D, _Container)

DOWN_TYPES = ("DownBlock2D", "AttnDownBlock2D")
UP_TYPES = ("UpBlock2D", "AttnUpBlock2D")


class AttentionBlock(_Container):
    """diffusers `Attention(..., residual_connection=True, bias=True, norm_num_groups=groups)` parameter container"""

    def __init__(self, channels, head_dim, groups, eps):
        super().__init__()
        self.dim_head = head_dim if head_dim is not None else channels
        self.heads = channels // self.dim_head
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])


class _Down(_Container):
    def __init__(self, in_ch, out_ch, temb_ch, num_layers, add_downsample, groups, eps, head_dim, attn):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups, eps) for i in range(num_layers)])
        self.attentions = (nn.ModuleList([AttentionBlock(out_ch, head_dim, groups, eps) for _ in range(num_layers)])
                           if attn else None)
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None


class _Mid(_Container):
    def __init__(self, ch, temb_ch, groups, eps, head_dim, add_attention):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([AttentionBlock(ch, head_dim, groups, eps) if add_attention else None])


class _Up(_Container):
    def __init__(self, in_ch, out_ch, prev_ch, temb_ch, num_layers, add_upsample, groups, eps, head_dim, attn):
        super().__init__()
        res = []
        for i in range(num_layers):
            skip = in_ch if i == num_layers - 1 else out_ch
            rin = prev_ch if i == 0 else out_ch
            res.append(ResnetBlock2D(rin + skip, out_ch, temb_ch, groups, eps))
        self.resnets = nn.ModuleList(res)
        self.attentions = (nn.ModuleList([AttentionBlock(out_ch, head_dim, groups, eps) for _ in range(num_layers)])
                           if attn else None)
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None


class DiffusersUNet2DWrapper(DiffusersUNet2DCondWrapper):
    """Constructor keywords and defaults of diffusers `UNet2DModel` (the reference passes *args / **kwargs through)."""

    def __init__(self, sample_size=None, in_channels=3, out_channels=3, center_input_sample=False,
                 time_embedding_type="positional", freq_shift=0, flip_sin_to_cos=True,
                 down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
                 up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels=(224, 448, 672, 896), layers_per_block=2, mid_block_scale_factor=1,
                 downsample_padding=1, downsample_type="conv", upsample_type="conv", dropout=0.0, act_fn="silu",
                 attention_head_dim: Optional[int] = 8, norm_num_groups=32, attn_norm_num_groups=None, norm_eps=1e-5,
                 resnet_time_scale_shift="default", add_attention=True, class_embed_type=None, num_class_embeds=None,
                 num_train_timesteps=None, **unused):
        nn.Module.__init__(self)
        if (time_embedding_type != "positional" or act_fn != "silu" or resnet_time_scale_shift != "default"
                or downsample_type != "conv" or upsample_type != "conv" or class_embed_type is not None
                or attn_norm_num_groups is not None or mid_block_scale_factor != 1 or downsample_padding != 1):
            raise NotImplementedError("built: UNet2DModel with positional time embedding, conv down / up-sampling, "
                                      "default ResNet blocks and nn.Embedding class conditioning")
        if not flip_sin_to_cos or freq_shift != 0:
            raise NotImplementedError("timestep embedding kernel implements flip_sin_to_cos=True, freq_shift=0")
        boc = list(block_out_channels)
        n = len(boc)
        temb_ch = boc[0] * 4
        self.in_channels, self.out_channels = in_channels, out_channels
        self.time_dim = boc[0]
        self.center_input_sample = center_input_sample
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.class_embedding = nn.Embedding(num_class_embeds, temb_ch) if num_class_embeds is not None else None
        hd = lambda ch: attention_head_dim if attention_head_dim is not None else ch
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            if t not in DOWN_TYPES:
                raise NotImplementedError(t)
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(_Down(in_ch, out_ch, temb_ch, layers_per_block, i != n - 1, norm_num_groups,
                                          norm_eps, hd(out_ch), t == "AttnDownBlock2D"))
        self.mid_block = _Mid(boc[-1], temb_ch, norm_num_groups, norm_eps, hd(boc[-1]), add_attention)
        self.up_blocks = nn.ModuleList()
        rboc = boc[::-1]
        out_ch = rboc[0]
        for i, t in enumerate(up_block_types):
            if t not in UP_TYPES:
                raise NotImplementedError(t)
            prev_ch, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, n - 1)]
            self.up_blocks.append(_Up(in_ch, out_ch, prev_ch, temb_ch, layers_per_block + 1, i != n - 1,
                                      norm_num_groups, norm_eps, hd(out_ch), t == "AttnUpBlock2D"))
        groups_out = norm_num_groups if norm_num_groups is not None else min(boc[0] // 4, 32)
        self.conv_norm_out = nn.GroupNorm(groups_out, boc[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.__dict__["_packs"] = {}

    supports_kv_cache = False

    def add_adapter(self, lora_config):
        raise NotImplementedError("LoRA adapters are built for the conditional denoisers of the distillation path")

    # ------------------------------------------------------------------------------------ engine pieces
    def _temb(self, timestep, class_labels, B, device):
        """SiLU(time_embedding(sinusoid(t)) [+ class_embedding[labels]]) as bf16 rows (every consumer applies SiLU)"""
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([float(timestep)], dtype=torch.float32, device=device)
        timestep = timestep.to(device=device, dtype=torch.float32).reshape(-1)
        if timestep.numel() == 1 and B > 1:
            timestep = timestep.expand(B)
        te = self.time_embedding
        p1 = self._pack("te1", lambda: LinearPack(te.linear_1)).pack()
        p2 = self._pack("te2", lambda: LinearPack(te.linear_2)).pack()
        t_emb = raw.timestep_embedding(timestep.contiguous(), self.time_dim)
        h = raw.silu_f32_to_bf16(raw.gemm(t_emb, p1["w"], bias=p1["b"], out_fp32=True))
        emb = raw.gemm(h, p2["w"], bias=p2["b"], out_fp32=True)
        if self.class_embedding is not None:
            if class_labels is None:
                raise ValueError("class_labels should be provided when doing class conditioning")
            rows = self.class_embedding.weight.detach().float()[class_labels.to(device).long().reshape(-1)]
            emb = emb + rows
        return raw.silu_f32_to_bf16(emb.contiguous())

    def _transformer(self, a, x, geom, ctx=None):
        """the spatial self-attention block that stands where the conditional UNet has a Transformer2DModel"""
        if a is None:
            return x
        B = geom[0]
        H, d = a.heads, a.dim_head
        dp = (d + 15) // 16 * 16
        hp = (H, d, dp) if dp != d else None
        inner = H * dp
        h = ops.group_norm(x, geom, a.group_norm, silu=False)
        qkv = ops.linear(h, self._pack(("qkv", id(a)), lambda: LinearPack([a.to_q, a.to_k, a.to_v], head_pad=hp)))
        o = ops.attention_self(qkv.view(B, -1, 3 * inner), H, head_dim=dp, scale=d ** -0.5).view(-1, inner)
        return ops.linear(o, self._pack(("o", id(a)), lambda: LinearPack(a.to_out[0], head_pad=hp, pad_cols=True)),
                          residual=x)

    # ------------------------------------------------------------------------------------ forward
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                conditioning: Dict[str, torch.Tensor] = None, *args, **kwargs):
        if not sample.is_cuda:
            raise RuntimeError("DiffusersUNet2DWrapper runs only on CUDA (B200) tensors: there is no CPU fallback")
        cond = conditioning["cond"] if conditioning is not None else {}       # reference wrapper :33-45
        self.__dict__["_kv_mode"] = None
        self.__dict__["_arena"] = ops.StatsArena(sample.device)
        try:
            return self._forward(sample, timestep, {"cond": {"vector": cond.get("vector"), "concat": cond.get("concat")}})
        finally:
            self.__dict__["_arena"] = None
