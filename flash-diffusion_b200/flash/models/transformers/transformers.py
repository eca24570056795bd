"""B200-native `DiffusersTransformer2DWrapper` — PixArt-alpha DiT (reference src/flash/models/transformers/tranformers.py:9-100
with the custom `AdaLayerNormSingle`, src/flash/models/transformers/utils.py:8-102; constructor kwargs as at
examples/train_flash_pixart.py:65-86).

Forward and backward on the hand-written kernels: activation gradients (incl. the input gradient of the frozen GAN
backbone) and LoRA gradients for every nn.Linear target of examples/train_flash_pixart.py:237-256 (attention with
head-padded packs, feed-forward, caption projection, AdaLN-single and its embedders) and for the 2x2 patch convolution,
which the "proj" target also names (peft `lora.Conv2d`).

Kernel mapping (UPSTREAM diffusers math, restated in oracle/dit.py):
  PatchEmbed conv 2x2/2    space-to-depth + 4-tap implicit-GEMM (fd_gemm conv mode) with the sin-cos position table added
                           as the residual of the same epilogue
  adaLN-single MLPs        small fd_gemm launches (M = batch)
  caption projection       fd_gemm with the gelu-tanh epilogue, then fd_gemm
  LN * (1+scale) + shift   fd_layernorm_modulate
  attention (d = 72)       q/k/v packs zero-padded to 80 channels per head -> fd_attn_fwd_generic / fd_attn_bwd_generic
                           (T5 key-padding mask as per-sample valid lengths)
  gate * f(x) + x          fd_gemm epilogue: bias -> (gelu-tanh) -> per-sample gate vector -> residual
  un-patchify              fd_unpatchify (keeps the first `in_channels` channels, as the reference slices them)
"""
from typing import Dict, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from ...b200 import ops, raw
from ...b200.ops import LinearPack, cache_of
from ..lora import inject_lora
from ..unets.unet import TimestepEmbedding, _Container


def _sincos_1d(embed_dim, pos):
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim, grid_size, base_size, interpolation_scale):
    """UPSTREAM diffusers `get_2d_sincos_pos_embed` (PixArt position table; not a parameter, not in the state dict)."""
    gh, gw = (grid_size, grid_size) if isinstance(grid_size, int) else grid_size
    g_h = np.arange(gh, dtype=np.float32) / (gh / base_size) / interpolation_scale
    g_w = np.arange(gw, dtype=np.float32) / (gw / base_size) / interpolation_scale
    # upstream builds np.meshgrid(g_w, g_h) and embeds the two coordinate planes (w first); the table is separable —
    # row i * gw + j = [sincos(g_w[j]) | sincos(g_h[i])] — so two 1-D tables are broadcast into the float32 result
    # (same float64 arithmetic per entry; the SD3 table is 36864 x 1536)
    e_w = _sincos_1d(embed_dim // 2, g_w).astype(np.float32)
    e_h = _sincos_1d(embed_dim // 2, g_h).astype(np.float32)
    out = np.empty((gh, gw, 2 * e_w.shape[1]), dtype=np.float32)
    out[:, :, :e_w.shape[1]] = e_w[None, :, :]
    out[:, :, e_w.shape[1]:] = e_h[:, None, :]
    return out.reshape(gh * gw, -1)


class PatchEmbed(_Container):
    def __init__(self, sample_size, patch_size, in_channels, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, patch_size, stride=patch_size)
        grid = sample_size // patch_size
        self.grid, self.base_size, self.interpolation_scale = grid, grid, max(sample_size // 64, 1)
        pe = sincos_2d(embed_dim, grid, base_size=grid, interpolation_scale=self.interpolation_scale)
        self.register_buffer("pos_embed", torch.from_numpy(pe).float()[None], persistent=False)

    def table(self, hh, ww):
        """Position table of an hh x ww patch grid.  diffusers' PatchEmbed.forward RECOMPUTES the 2-D sincos table
        for (height, width) whenever the input grid differs from the configured square one (same base_size and
        interpolation_scale) — slicing the first rows of the square table would give wrong 2-D positions."""
        if (hh, ww) == (self.grid, self.grid):
            return self.pos_embed[0]
        pe = sincos_2d(self.pos_embed.shape[-1], (hh, ww), base_size=self.base_size,
                       interpolation_scale=self.interpolation_scale)
        return torch.from_numpy(pe).float()


class AdaLayerNormSingle(_Container):
    """reference src/flash/models/transformers/utils.py:8-102 (same attribute / key names)."""

    def __init__(self, time_embed_dim, timesteps_embedding_num_channels=256, projection_class_embeddings_input_dim=None,
                 use_concat_conditioning=False, num_vector_conditionings=None):
        super().__init__()
        self.num_channels = timesteps_embedding_num_channels
        self.timestep_embedder = TimestepEmbedding(timesteps_embedding_num_channels, time_embed_dim)
        self.projection_class_embeddings_input_dim = projection_class_embeddings_input_dim
        self.use_concat_conditioning = use_concat_conditioning
        self.num_vector_conditionings = num_vector_conditionings
        if projection_class_embeddings_input_dim is not None:
            if not use_concat_conditioning:
                self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, time_embed_dim)
            else:
                assert num_vector_conditionings is not None, \
                    "num_vector_conditionings must be provided if use_concat_conditioning is True"
                self.add_embedding = nn.ModuleList(
                    [TimestepEmbedding(projection_class_embeddings_input_dim, time_embed_dim // num_vector_conditionings)
                     for _ in range(num_vector_conditionings)])
        self.linear = nn.Linear(time_embed_dim, 6 * time_embed_dim, bias=True)


class TextProjection(_Container):
    def __init__(self, in_features, hidden):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden)
        self.linear_2 = nn.Linear(hidden, hidden)


class Attention(_Container):
    def __init__(self, dim, cross_dim, heads, dim_head, bias):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head, self.is_cross = heads, dim_head, cross_dim is not None
        self.to_q = nn.Linear(dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_dim or dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_dim or dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])


class GELUProj(_Container):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)


class FeedForward(_Container):
    def __init__(self, dim, mult=4, geglu=False):
        super().__init__()
        self.geglu = geglu               # diffusers GEGLU: proj to 2 * inner, value * gelu(gate)
        self.net = nn.ModuleList([GELUProj(dim, dim * mult * (2 if geglu else 1)), nn.Dropout(0.0),
                                  nn.Linear(dim * mult, dim)])


class AdaBlock(_Container):
    """diffusers `BasicTransformerBlock` with norm_type="ada_norm_single": norm1 / norm2 carry parameters only when
    `norm_elementwise_affine` (PixArt: False, so the state dict has none); attn2 is a cross-attention, a second
    self-attention (`double_self_attention`) or absent (no cross_attention_dim); no norm3."""

    def __init__(self, dim, heads, dim_head, cross_dim, bias, affine=False, eps=1e-6, geglu=False,
                 double_self_attention=False):
        super().__init__()
        self.scale_shift_table = nn.Parameter(torch.randn(6, dim) / dim ** 0.5)
        if affine:
            self.norm1 = nn.LayerNorm(dim, eps=eps, elementwise_affine=True)
        self.attn1 = Attention(dim, None, heads, dim_head, bias)
        self.attn2 = None
        if cross_dim is not None or double_self_attention:
            if affine:
                self.norm2 = nn.LayerNorm(dim, eps=eps, elementwise_affine=True)
            self.attn2 = Attention(dim, None if double_self_attention else cross_dim, heads, dim_head, bias)
        elif affine:
            self.norm2 = nn.LayerNorm(dim, eps=eps, elementwise_affine=True)       # still modulates the feed-forward
        self.ff = FeedForward(dim, geglu=geglu)


class DiffusersTransformer2DWrapper(nn.Module):
    def __init__(self, time_embed_dim: int = 256, timesteps_embedding_num_channels: int = 256,
                 projection_class_embeddings_input_dim: Optional[int] = None,
                 use_concat_vector_conditioning: bool = False, num_vector_conditionings: Optional[int] = None,
                 sample_size=None, num_layers=1, attention_head_dim=88, in_channels=None, out_channels=None,
                 patch_size=None, attention_bias=False, num_attention_heads=16, cross_attention_dim=None,
                 activation_fn="geglu", norm_type="layer_norm", norm_elementwise_affine=True,
                 norm_eps=1e-5, caption_channels=None, double_self_attention=False, **unused):
        """Keyword defaults are diffusers' `Transformer2DModel` defaults (the reference's wrapper subclasses it,
        transformers/tranformers.py:19-47), so a caller that omits e.g. `caption_channels` gets what the reference
        builds; the PixArt-alpha values are spelled out by the example script (examples/train_flash_pixart.py:65-86)."""
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        if norm_type != "ada_norm_single" or activation_fn not in ("gelu-approximate", "geglu") or patch_size != 2 \
                or sample_size is None or in_channels is None:
            raise NotImplementedError("built: Transformer2DModel with norm_type='ada_norm_single', patch_size 2 and a "
                                      "'gelu-approximate' (PixArt-alpha) or 'geglu' feed-forward")
        D = num_attention_heads * attention_head_dim
        if time_embed_dim != D:
            raise ValueError(f"time_embed_dim ({time_embed_dim}) must equal the inner dimension ({D}): the adaLN-single "
                             "table is added to 6 * inner_dim modulation rows")
        self.patch_size, self.out_channels, self.in_channels, self.norm_eps = patch_size, out_channels, in_channels, norm_eps
        self.inner_dim, self.sample_size = D, sample_size
        self.pos_embed = PatchEmbed(sample_size, patch_size, in_channels, D)
        self.adaln_single = AdaLayerNormSingle(time_embed_dim, timesteps_embedding_num_channels,
                                               projection_class_embeddings_input_dim, use_concat_vector_conditioning,
                                               num_vector_conditionings)
        if caption_channels is not None:      # PixArt; without it the text states feed attn2.to_k / to_v directly
            self.caption_projection = TextProjection(caption_channels, D)
        else:
            self.caption_projection = None
        self.transformer_blocks = nn.ModuleList(
            [AdaBlock(D, num_attention_heads, attention_head_dim, cross_attention_dim, attention_bias,
                      affine=bool(norm_elementwise_affine), eps=norm_eps, geglu=activation_fn == "geglu",
                      double_self_attention=double_self_attention)
             for _ in range(num_layers)])
        self.scale_shift_table = nn.Parameter(torch.randn(2, D) / D ** 0.5)
        self.proj_out = nn.Linear(D, patch_size * patch_size * out_channels)
        self.__dict__["_packs"] = {}

    # ------------------------------------------------------------------------------------ helpers
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

    def freeze(self):
        self.eval()
        for p in self.parameters():
            p.requires_grad = False

    def add_adapter(self, lora_config):
        """LoRA on the targets' nn.Linear modules and, as peft does for the "proj" target, on the patch convolution."""
        inject_lora(self, lora_config)
        self.__dict__["_packs"] = {}
        return self

    @staticmethod
    def _silu_bf16(x):
        """silu(fp32) -> bf16 on a [B, D] embedding; with a gradient it is three torch ops on B rows."""
        if torch.is_grad_enabled() and x.requires_grad:
            return torch.nn.functional.silu(x).to(torch.bfloat16)
        return raw.silu_f32_to_bf16(x)

    def _mlp_rows(self, x_bf16, te: TimestepEmbedding, key):
        """TimestepEmbedding on [B, in] rows: linear_1 -> SiLU -> linear_2, fp32 out."""
        l1 = self._pack((key, 1), lambda: LinearPack(te.linear_1))
        l2 = self._pack((key, 2), lambda: LinearPack(te.linear_2))
        h = self._silu_bf16(ops.linear(x_bf16, l1, out_fp32=True))
        return ops.linear(h, l2, out_fp32=True)

    def _adaln(self, timestep, vector, B, dev):
        ad = self.adaln_single
        t_emb = raw.timestep_embedding(timestep, ad.num_channels)
        emb = self._mlp_rows(t_emb, ad.timestep_embedder, "te")
        if ad.projection_class_embeddings_input_dim is not None:
            if vector is None:
                raise ValueError("vector conditioning is required by this adaln_single configuration")
            v = raw.cast_scale(vector.detach().float().contiguous(), 1.0)
            if isinstance(ad.add_embedding, nn.ModuleList):
                n = ad.num_vector_conditionings
                w = v.shape[1] // n
                parts = [self._mlp_rows(v[:, i * w:(i + 1) * w].contiguous(), ad.add_embedding[i], ("ae", i))
                         for i in range(n)]
                emb = emb + torch.cat(parts, dim=1)
            else:
                emb = emb + self._mlp_rows(v, ad.add_embedding, "ae")
        t6 = ops.linear(self._silu_bf16(emb), self._pack("adaln_linear", lambda: LinearPack(ad.linear)), out_fp32=True)
        return t6, emb

    def _attention(self, a: Attention, x, ctx, B, kv_len, residual, gate=None, rows=0):
        H, d = a.heads, a.dim_head
        dp = (d + 15) // 16 * 16
        hp = (H, d, dp) if dp != d else None
        inner = H * dp
        if not a.is_cross:
            qkv = ops.linear(x, self._pack(("qkv", id(a)), lambda: LinearPack([a.to_q, a.to_k, a.to_v], head_pad=hp)))
            o = ops.attention_self(qkv.view(B, -1, 3 * inner), H, head_dim=dp, scale=d ** -0.5)
        else:
            q = ops.linear(x, self._pack(("q", id(a)), lambda: LinearPack(a.to_q, head_pad=hp)))
            kv = ops.linear(ctx, self._pack(("kv", id(a)), lambda: LinearPack([a.to_k, a.to_v], head_pad=hp)))
            o = ops.attention_cross(q.view(B, -1, inner), kv.view(B, -1, 2 * inner), H, head_dim=dp, scale=d ** -0.5,
                                    kv_len=kv_len)
        out_pack = self._pack(("o", id(a)), lambda: LinearPack(a.to_out[0], head_pad=hp, pad_cols=True))
        if gate is not None:
            return ops.gated_linear(o.view(-1, inner), out_pack, gate, residual, rows)
        return ops.linear(o.view(-1, inner), out_pack, residual=residual)

    # ------------------------------------------------------------------------------------ forward
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                conditioning: Dict[str, torch.Tensor], hidden_states_masks: Optional[torch.Tensor] = None,
                *args, **kwargs):
        assert isinstance(conditioning, dict), "conditionings must be a dictionary"
        if not sample.is_cuda:
            raise RuntimeError("DiffusersTransformer2DWrapper runs only on CUDA (B200) tensors: there is no CPU fallback")
        cond = conditioning["cond"]
        vector, crossattn, concat = cond.get("vector"), cond.get("crossattn"), cond.get("concat")
        mask = cond.get("attention_mask")
        c_keep = sample.shape[1]
        if concat is not None:
            sample = torch.cat([sample, concat], dim=1)
        B, Cin, H, W = sample.shape
        dev, p, D = sample.device, self.patch_size, self.inner_dim
        hh, ww = H // p, W // p
        N = hh * ww
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([float(timestep)], dtype=torch.float32, device=dev)
        timestep = timestep.detach().to(device=dev, dtype=torch.float32).reshape(-1)
        if timestep.numel() == 1 and B > 1:
            timestep = timestep.expand(B)
        timestep = timestep.contiguous()
        t6, emb = self._adaln(timestep, vector, B, dev)
        # caption projection (gelu-tanh in the epilogue of linear_1)
        cp = self.caption_projection
        ctx = kv_len = None
        needs_ctx = any(b.attn2 is not None and b.attn2.is_cross for b in self.transformer_blocks)
        if needs_ctx:
            if crossattn is None:
                raise ValueError("this configuration has cross-attention layers: conditioning['cond']['crossattn'] is required")
            T = crossattn.shape[1]
            ctx = raw.cast_scale(crossattn.detach().float().contiguous().view(B * T, -1), 1.0)
            if cp is not None:
                c1 = ops.linear(ctx, self._pack("cp1", lambda: LinearPack(cp.linear_1)), act=1)
                ctx = ops.linear(c1, self._pack("cp2", lambda: LinearPack(cp.linear_2)))
            if mask is not None:      # T5 padding mask (ones then zeros): per-sample number of valid keys
                kv_len = mask.to(device=dev).reshape(B, T).sum(dim=1).to(torch.int32).contiguous()
        # patch embedding: 2x2 stride-2 conv == 4-tap implicit GEMM over the space-to-depth image (+ position table)
        cpad = (Cin + 7) // 8 * 8
        pe = self.pos_embed

        def build_patch():
            wt = pe.proj.weight.detach().float()                      # [D, Cin, 2, 2]
            buf = torch.zeros((D, 4, 64), device=dev)
            buf[:, :, :Cin] = wt.permute(0, 2, 3, 1).reshape(D, 4, Cin)
            pos = pe.table(H // 2, W // 2).to(dev)
            return {"w": raw.cast_scale(buf.reshape(D, 256), 1.0),
                    "w_t": raw.cast_scale(buf.reshape(D, 256).t().contiguous(), 1.0),
                    "b": pe.proj.bias.detach().float().contiguous(), "pos": raw.cast_scale(pos.contiguous(), 1.0)}
        pk = cache_of(pe.proj).get(("patch", H // 2, W // 2), [pe.proj.weight, pe.proj.bias], build_patch)
        pos_b = self._pack(("pos_tiled", B, H // 2, W // 2), lambda: pk["pos"].repeat(B, 1).contiguous())
        h = patch_embed(sample.float(), pk, pos_b, (B, Cin, H, W, cpad), patch_lora_pack(pe.proj, Cin, dev))
        # AdaLN parameters of every block in one small op: [L, B, 6, D]
        tables = self._pack("tables", lambda: torch.stack([b.scale_shift_table.detach().float()
                                                           for b in self.transformer_blocks]))
        mods = (tables[:, None] + t6.view(1, B, 6, D)).contiguous()
        eps = self.norm_eps
        def affine_mod(norm, scale, shift):
            """(LN(x) * gamma + beta) * (1 + scale) + shift == LN(x) * (1 + scale') + shift' with
            scale' = gamma * (1 + scale) - 1, shift' = beta * (1 + scale) + shift  ([B, D] rows, host-side)"""
            if norm is None:
                return scale, shift
            g, b = norm.weight.detach().float(), norm.bias.detach().float()
            return (g * (1 + scale) - 1).contiguous(), (b * (1 + scale) + shift).contiguous()

        for li, blk in enumerate(self.transformer_blocks):
            m = mods[li]                                              # [B, 6, D]: shift/scale/gate msa, mlp
            sc, sh = affine_mod(getattr(blk, "norm1", None), m[:, 1], m[:, 0])
            n1 = ops.modulate(h, sc, sh, N, eps)
            h = self._attention(blk.attn1, n1, None, B, None, residual=h, gate=m[:, 2], rows=N)
            if blk.attn2 is not None:     # ada_norm_single feeds attn2 the un-normalised states (UPSTREAM forward)
                h = self._attention(blk.attn2, h, ctx if blk.attn2.is_cross else None, B, kv_len, residual=h)
            sc, sh = affine_mod(getattr(blk, "norm2", None), m[:, 4], m[:, 3])
            n2 = ops.modulate(h, sc, sh, N, eps)
            if blk.ff.geglu:
                f = ops.geglu(n2, self._pack(("ff1", id(blk)), lambda: LinearPack(blk.ff.net[0].proj, geglu=True)))
            else:
                f = ops.linear(n2, self._pack(("ff1", id(blk)), lambda: LinearPack(blk.ff.net[0].proj)), act=1)
            h = ops.gated_linear(f, self._pack(("ff2", id(blk)), lambda: LinearPack(blk.ff.net[2])), m[:, 5], h, N)
        fin = (self.scale_shift_table.detach().float()[None] + emb[:, None]).contiguous()     # [B, 2, D]
        nf = ops.modulate(h, fin[:, 1], fin[:, 0], N, 1e-6)           # norm_out: LayerNorm(eps=1e-6, no affine), UPSTREAM
        out = ops.linear(nf, self._pack("proj_out", lambda: LinearPack(self.proj_out)), out_fp32=True)
        return ops.unpatchify(out, B, hh, ww, p, self.out_channels, min(c_keep, self.out_channels))


def patch_lora_pack(proj, Cin, dev):
    """Kernel-side view of a LoRA-wrapped patch convolution (peft `lora.Conv2d`: lora_A is a 2x2/2 conv to r channels,
    lora_B a 1x1 conv): A as a [r, 4*64] matrix in the tap-major column order of the patch GEMM, s*B as [D, r]."""
    if not hasattr(proj, "lora_A"):
        return None

    def build():
        A = proj.lora_A["default"].weight.detach().float()                              # [r, Cin, 2, 2]
        Bm = proj.lora_B["default"].weight.detach().float()[:, :, 0, 0] * proj.scaling   # [D, r]
        r = A.shape[0]
        buf = torch.zeros((r, 4, 64), device=dev)
        buf[:, :, :Cin] = A.permute(0, 2, 3, 1).reshape(r, 4, Cin)
        a = buf.reshape(r, 256)
        return {"a": raw.cast_scale(a.contiguous(), 1.0), "a_t": raw.cast_scale(a.t().contiguous(), 1.0),
                "b": raw.cast_scale(Bm.contiguous(), 1.0), "b_t": raw.cast_scale(Bm.t().contiguous(), 1.0),
                "r": r, "scaling": proj.scaling, "Cin": Cin}
    pack = cache_of(proj).get(("patch_lora", Cin), [proj.lora_A["default"].weight, proj.lora_B["default"].weight], build)
    return dict(pack, params=(proj.lora_A["default"].weight, proj.lora_B["default"].weight))


def _patch_embed_fwd(sample, pk, pos_b, geom, lora=None):
    """2x2/2 patch convolution + position table: space-to-depth, then a 4-tap implicit GEMM whose taps are the four
    phase images; the table rides in the epilogue as the residual, a LoRA on the convolution as a second K segment.
    Returns (tokens, space-to-depth image, x A^T or None)."""
    B, Cin, H, W, cpad = geom
    hh, ww = H // 2, W // 2
    x = raw.space_to_depth(raw.nchw_to_nhwc(sample, cpad).view(B * H * W, cpad), B, H, W, cpad)
    conv = dict(NB_in=4 * B, H=hh, W=ww, C=cpad, taps=[(ph * B, 0, 0) for ph in range(4)])
    t = a2 = b2 = None
    if lora is not None:
        t = raw.gemm(x, lora["a"], M=B * hh * ww, conv=conv)
        a2, b2 = t, lora["b"]
    return raw.gemm(x, pk["w"], a2=a2, b2=b2, bias=pk["b"], residual=pos_b, M=B * hh * ww, conv=conv), x, t


class _PatchEmbedFn(torch.autograd.Function):
    """Gradients of the patch embedding: with respect to its input (the GAN generator turn differentiates the frozen
    backbone with respect to its input, reference flash_diffusion_model.py:563-592) — d(tokens) W gives the four phase
    images, depth-to-space and NHWC->NCHW undo the forward re-layout — and with respect to a LoRA on the convolution."""

    @staticmethod
    def forward(ctx, sample, pk, pos_b, geom, lora, *lora_params):
        h, x, t = _patch_embed_fwd(sample, pk, pos_b, geom, lora)
        ctx.pk, ctx.geom, ctx.lora = pk, geom, lora
        ctx.save_for_backward(x if lora is not None else None, t)
        return h

    @staticmethod
    def backward(ctx, dh):
        B, Cin, H, W, cpad = ctx.geom
        hh, ww = H // 2, W // 2
        BN = B * hh * ww
        lora = ctx.lora
        x, t = ctx.saved_tensors
        dh = dh.contiguous()
        grads = ()
        dt = None
        if lora is not None:
            r = lora["r"]
            dt = raw.gemm(dh, lora["b_t"])                                               # [BN, r] = dh (sB)
            X = torch.zeros((BN, 4, 64), device=dh.device, dtype=dh.dtype)                # the patch GEMM's A matrix
            X[:, :, :cpad] = x.view(4, BN, cpad).permute(1, 0, 2)
            d_a = raw.gemm(raw.transpose(dt), raw.transpose(X.view(BN, 256)), out_fp32=True)      # [r, 256]
            d_b = raw.gemm(raw.transpose(dh), raw.transpose(t), out_fp32=True)                    # [D, r]
            g_a = d_a.view(r, 4, 64)[:, :, :Cin].permute(0, 2, 1).reshape(r, Cin, 2, 2)
            grads = (g_a, (d_b * lora["scaling"]).view(d_b.shape[0], r, 1, 1))
        dx = None
        if ctx.needs_input_grad[0]:
            d = raw.gemm(dh, ctx.pk["w_t"], a2=dt, b2=lora["a_t"] if lora is not None else None)   # [BN, 4*64]
            phases = d.view(BN, 4, 64)[:, :, :cpad].permute(1, 0, 2).contiguous().view(4 * BN, cpad)
            dx = raw.nhwc_to_nchw(raw.depth_to_space(phases, B, H, W, cpad), B, Cin, H, W)
        return (dx, None, None, None, None, *grads)


def patch_embed(sample, pk, pos_b, geom, lora=None):
    """sample NCHW fp32 -> tokens [B*N, D] bf16; pk = {"w" [D,256], "w_t", "b", "pos"} (see the wrappers), lora from
    patch_lora_pack."""
    params = lora["params"] if lora is not None else ()
    if torch.is_grad_enabled() and (sample.requires_grad or any(q.requires_grad for q in params)):
        return _PatchEmbedFn.apply(sample, pk, pos_b, geom, lora, *params)
    return _patch_embed_fwd(sample, pk, pos_b, geom, lora)[0]
