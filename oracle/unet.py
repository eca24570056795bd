"""ORACLE (test infrastructure — never imported by the product path).

Pure-PyTorch fp32 restatement of the UPSTREAM arithmetic behind the reference's denoiser
wrapper `DiffusersUNet2DCondWrapper` (reference: src/flash/models/unets/unet.py:55-127), i.e.
diffusers' `UNet2DConditionModel` with the constructor kwargs the reference passes
(examples/train_flash_sdxl.py:66-118, examples/train_flash_sd.py:56-114,
tests/test_flash/test_flash_diffusion.py:44-60) and peft-0.9 LoRA injection
(examples/train_flash_sdxl.py:210-217).

PARITY UNPINNED: diffusers (fork `initml/diffusers@clement/feature/flash`, requirements.txt:1) and
peft 0.9.0 (setup.py:37) are not installable here and the reference holds no golden vectors
(SURVEY.md §8c), so this file restates their published module math
and is pinned only structurally: state-dict key names (src/flash/trainer/utils.py:58-61,195,
examples/train_flash_sdxl.py:123-134, examples/train_flash_sd.py:119-152) and parameter counts
(SURVEY.md Appendix B.6).
"""
import math
from typing import Dict, List, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, freq_shift: float = 0.0):
    """diffusers `Timesteps` / get_timestep_embedding (max_period 10000, scale 1)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device)
    exponent = exponent / (half - freq_shift)
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class LoRALinear(nn.Module):
    """peft 0.9 `lora.Linear`: y = base(x) + lora_B(lora_A(x)) * (alpha / r)."""

    def __init__(self, base: nn.Linear, r: int, lora_alpha: int, init: Union[bool, str] = True):
        super().__init__()
        self.base_layer = base
        self.r = r
        self.scaling = lora_alpha / r
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        if init == "gaussian":
            nn.init.normal_(self.lora_A["default"].weight, std=1.0 / r)
        else:
            nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)

    def forward(self, x):
        return self.base_layer(x) + self.lora_B["default"](self.lora_A["default"](x)) * self.scaling


class LoRAConv2d(nn.Module):
    """peft 0.9 `lora.Conv2d`: y = base(x) + lora_B(lora_A(x)) * (alpha / r); lora_A has the base kernel / stride /
    padding, lora_B is 1x1.  Reached by the DiT recipes, whose target "proj" also names the patch convolution
    (examples/train_flash_pixart.py:239-252, examples/train_flash_sd3.py:104-117)."""

    def __init__(self, base: nn.Conv2d, r: int, lora_alpha: int, init: Union[bool, str] = True):
        super().__init__()
        self.base_layer = base
        self.r = r
        self.scaling = lora_alpha / r
        self.lora_A = nn.ModuleDict({"default": nn.Conv2d(base.in_channels, r, base.kernel_size, base.stride,
                                                         base.padding, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Conv2d(r, base.out_channels, 1, 1, bias=False)})
        if init == "gaussian":
            nn.init.normal_(self.lora_A["default"].weight, std=1.0 / r)
        else:
            nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)

    def forward(self, x):
        return self.base_layer(x) + self.lora_B["default"](self.lora_A["default"](x)) * self.scaling


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        context = x if context is None else context
        B, N, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(context), self.to_v(context)
        d = q.shape[-1] // self.heads
        q = q.view(B, N, self.heads, d).transpose(1, 2)
        k = k.view(B, -1, self.heads, d).transpose(1, 2)
        v = v.view(B, -1, self.heads, d).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
        o = torch.softmax(s, dim=-1) @ v
        o = o.transpose(1, 2).reshape(B, N, self.heads * d)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), context)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, num_layers, cross_attention_dim, groups=32,
                 use_linear_projection=False):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner)
            self.proj_out = nn.Linear(inner, in_channels)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner, 1)
            self.proj_out = nn.Conv2d(inner, in_channels, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])

    def forward(self, x, context):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x)
        if self.use_linear_projection:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(B, H * W, -1)
        for blk in self.transformer_blocks:
            h = blk(h, context)
        if self.use_linear_projection:
            h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.reshape(B, H, W, -1).permute(0, 3, 1, 2))
        return h + res


class Downsample2D(nn.Module):
    def __init__(self, channels, padding=1):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """DownBlock2D / CrossAttnDownBlock2D."""

    def __init__(self, in_ch, out_ch, temb_ch, num_layers, add_downsample, groups, eps, attn=None):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups, eps) for i in range(num_layers)])
        if attn is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(in_channels=out_ch, **attn) for _ in range(num_layers)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, x, temb, context):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    """UNetMidBlock2DCrossAttn."""

    def __init__(self, ch, temb_ch, groups, eps, attn):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(in_channels=ch, **attn)])

    def forward(self, x, temb, context):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, context)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    """UpBlock2D / CrossAttnUpBlock2D."""

    def __init__(self, in_ch, out_ch, prev_ch, temb_ch, num_layers, add_upsample, groups, eps, attn=None):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_ch if i == num_layers - 1 else out_ch
            rin = prev_ch if i == 0 else out_ch
            resnets.append(ResnetBlock2D(rin + skip, out_ch, temb_ch, groups, eps))
        self.resnets = nn.ModuleList(resnets)
        if attn is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(in_channels=out_ch, **attn) for _ in range(num_layers)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, x, skips: List[torch.Tensor], temb, context):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


def _per_block(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


class UNet2DConditionOracle(nn.Module):
    """Same constructor kwargs / state-dict keys / forward contract as the reference's
    `DiffusersUNet2DCondWrapper` (src/flash/models/unets/unet.py:55-127)."""

    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                 mid_block_type="UNetMidBlock2DCrossAttn",
                 up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, dropout=0.0, act_fn="silu", norm_num_groups=32,
                 norm_eps=1e-5, cross_attention_dim=1280, transformer_layers_per_block=1,
                 reverse_transformer_layers_per_block=None, attention_head_dim=8, num_attention_heads=None,
                 use_linear_projection=False, class_embed_type=None, projection_class_embeddings_input_dim=None,
                 **unused):
        super().__init__()
        assert mid_block_type == "UNetMidBlock2DCrossAttn" and act_fn == "silu"
        n = len(down_block_types)
        boc = list(block_out_channels)
        heads = _per_block(num_attention_heads or attention_head_dim, n)   # diffusers naming quirk
        tlpb = _per_block(transformer_layers_per_block, n)
        lpb = _per_block(layers_per_block, n)
        temb_ch = boc[0] * 4
        self.cfg = dict(in_channels=in_channels, out_channels=out_channels, flip_sin_to_cos=flip_sin_to_cos,
                        freq_shift=freq_shift, time_dim=boc[0], center_input_sample=center_input_sample)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        if class_embed_type == "projection":
            self.class_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, temb_ch)
        else:
            assert class_embed_type is None, class_embed_type
            self.class_embedding = None

        def attn_kwargs(i, ch):
            return dict(heads=heads[i], dim_head=ch // heads[i], num_layers=tlpb[i],
                        cross_attention_dim=cross_attention_dim, groups=norm_num_groups,
                        use_linear_projection=use_linear_projection)

        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            attn = attn_kwargs(i, out_ch) if t == "CrossAttnDownBlock2D" else None
            assert t in ("CrossAttnDownBlock2D", "DownBlock2D"), t
            self.down_blocks.append(DownBlock(in_ch, out_ch, temb_ch, lpb[i], i != n - 1, norm_num_groups, norm_eps, attn))
        self.mid_block = MidBlock(boc[-1], temb_ch, norm_num_groups, norm_eps, attn_kwargs(n - 1, boc[-1]))
        self.up_blocks = nn.ModuleList()
        rboc = boc[::-1]
        rheads, rtl, rlpb = heads[::-1], tlpb[::-1], lpb[::-1]
        if reverse_transformer_layers_per_block is not None:
            rtl = _per_block(reverse_transformer_layers_per_block, n)
        out_ch = rboc[0]
        for i, t in enumerate(up_block_types):
            prev_ch, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, n - 1)]
            assert t in ("CrossAttnUpBlock2D", "UpBlock2D"), t
            attn = None
            if t == "CrossAttnUpBlock2D":
                attn = dict(heads=rheads[i], dim_head=out_ch // rheads[i], num_layers=rtl[i],
                            cross_attention_dim=cross_attention_dim, groups=norm_num_groups,
                            use_linear_projection=use_linear_projection)
            self.up_blocks.append(UpBlock(in_ch, out_ch, prev_ch, temb_ch, rlpb[i] + 1, i != n - 1,
                                          norm_num_groups, norm_eps, attn))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    # ------------------------------------------------------------------ reference wrapper API
    def forward(self, sample, timestep, conditioning: Dict[str, Dict[str, torch.Tensor]],
                down_intrablock_additional_residuals=None, return_intermediate=False, *args, **kwargs):
        assert isinstance(conditioning, dict), "conditionings must be a dictionary"
        assert down_intrablock_additional_residuals is None, "T2I adapter residuals are out of scope"
        cond = conditioning["cond"]
        class_labels, context, concat = cond.get("vector"), cond.get("crossattn"), cond.get("concat")
        if concat is not None:
            sample = torch.cat([sample, concat], dim=1)
        if self.cfg["center_input_sample"]:
            sample = 2 * sample - 1.0
        B = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        elif timestep.dim() == 0:
            timestep = timestep[None].to(sample.device)
        timestep = timestep.expand(B)
        t_emb = timestep_embedding(timestep, self.cfg["time_dim"], self.cfg["flip_sin_to_cos"], self.cfg["freq_shift"])
        emb = self.time_embedding(t_emb.to(sample.dtype))
        if self.class_embedding is not None:
            assert class_labels is not None, "class_labels should be provided when num_class_embeds > 0"
            emb = emb + self.class_embedding(class_labels.to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, context)
            skips.extend(outs)
        x = self.mid_block(x, emb, context)
        if return_intermediate:     # fork-only kwarg; decision (2) of SURVEY §8c: early exit after the mid block
            return x
        for blk in self.up_blocks:
            x = blk(x, skips, emb, context)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return x

    def freeze(self):
        self.eval()
        for p in self.parameters():
            p.requires_grad = False

    def add_adapter(self, lora_config):
        """diffusers PeftAdapterMixin.add_adapter -> peft inject_adapter_in_model: wrap every Linear / Conv2d whose
        name ends with one of target_modules; freeze everything that is not a LoRA weight."""
        targets = list(lora_config.target_modules)
        for p in self.parameters():
            p.requires_grad = False
        for name, module in list(self.named_modules()):
            for child_name, child in list(module.named_children()):
                full = f"{name}.{child_name}" if name else child_name
                if isinstance(child, (nn.Linear, nn.Conv2d)) and any(full == t or full.endswith("." + t) for t in targets):
                    cls = LoRALinear if isinstance(child, nn.Linear) else LoRAConv2d
                    lora = cls(child, lora_config.r, lora_config.lora_alpha, lora_config.init_lora_weights)
                    if isinstance(module, nn.ModuleList):
                        module[int(child_name)] = lora
                    else:
                        setattr(module, child_name, lora)
        return self


class LoraConfig:
    """The peft.LoraConfig fields the reference uses (examples/train_flash_sdxl.py:210-216)."""

    def __init__(self, r=8, lora_alpha=8, init_lora_weights=True, target_modules=None, **unused):
        self.r, self.lora_alpha, self.init_lora_weights = r, lora_alpha, init_lora_weights
        self.target_modules = target_modules or []


SDXL_KWARGS = dict(
    in_channels=4, out_channels=4, down_block_types=["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"],
    up_block_types=["CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"], block_out_channels=[320, 640, 1280],
    layers_per_block=2, cross_attention_dim=2048, transformer_layers_per_block=[1, 2, 10],
    attention_head_dim=[5, 10, 20], use_linear_projection=True, class_embed_type="projection",
    projection_class_embeddings_input_dim=2816)   # examples/train_flash_sdxl.py:66-118

SD15_KWARGS = dict(
    in_channels=4, out_channels=4,
    down_block_types=["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
    up_block_types=["UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"],
    block_out_channels=[320, 640, 1280, 1280], layers_per_block=2, cross_attention_dim=768,
    transformer_layers_per_block=1, attention_head_dim=8, use_linear_projection=True,
    class_embed_type=None)                          # examples/train_flash_sd.py:56-114
