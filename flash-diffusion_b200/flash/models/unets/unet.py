"""B200-native `DiffusersUNet2DCondWrapper` (drop-in for reference src/flash/models/unets/unet.py:55-127).

The reference subclasses diffusers' `UNet2DConditionModel`; here the same constructor kwargs
(examples/train_flash_sdxl.py:66-118, examples/train_flash_sd.py:56-114) build the same module tree
with the same state-dict keys, but the modules are parameter containers only: `forward` walks the tree
and launches the hand-written sm_100a kernels of libflashb200.so (flash.b200.ops) on channels-last bf16
activations.  There is no eager/PyTorch/CPU fallback: a CPU tensor or a missing library raises.

Math (UPSTREAM diffusers, restated in SURVEY.md §8a-L1 and oracle/unet.py):
  ResnetBlock2D        GN32+SiLU -> conv3x3 (+time-embedding row vector) -> GN32+SiLU -> conv3x3 (+1x1 shortcut
                       accumulated as a second K segment of the same GEMM, or residual add in the epilogue)
  Transformer2DModel   GN32(eps 1e-6) -> proj_in -> N x BasicTransformerBlock -> proj_out (+residual epilogue)
  BasicTransformerBlock  LN -> fused QKV GEMM (LoRA folded as K-segment 2) -> FlashAttention kernel -> out-proj
                       (+residual) ; LN -> q / fused kv GEMMs -> cross-attention -> out-proj ; LN -> GEGLU GEMM
                       (activation in the epilogue) -> GEMM (+residual)
"""
from typing import Dict, Union

import torch
import torch.nn as nn

from ...b200 import ops, raw
from ...b200.ops import ConvPack, LinearPack
from ..lora import LoRALinear, inject_lora


def _per_block(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the B200 engine in DiffusersUNet2DCondWrapper.forward runs it")


class TimestepEmbedding(_Container):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class ResnetBlock2D(_Container):
    def __init__(self, in_channels, out_channels, temb_channels, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class Attention(_Container):
    def __init__(self, query_dim, cross_attention_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.is_cross = cross_attention_dim is not None
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])


class GEGLU(_Container):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_Container):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])


class BasicTransformerBlock(_Container):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)


class Transformer2DModel(_Container):
    def __init__(self, heads, dim_head, in_channels, num_layers, cross_attention_dim, groups, use_linear_projection):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        if use_linear_projection:
            self.proj_in, self.proj_out = nn.Linear(in_channels, inner), nn.Linear(inner, in_channels)
        else:
            self.proj_in, self.proj_out = nn.Conv2d(in_channels, inner, 1), nn.Conv2d(inner, in_channels, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])


class Downsample2D(_Container):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)


class Upsample2D(_Container):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)


class DownBlock(_Container):
    def __init__(self, in_ch, out_ch, temb_ch, num_layers, add_downsample, groups, eps, attn=None):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups, eps) for i in range(num_layers)])
        self.attentions = (nn.ModuleList([Transformer2DModel(in_channels=out_ch, **attn) for _ in range(num_layers)])
                           if attn is not None else None)
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None


class MidBlock(_Container):
    def __init__(self, ch, temb_ch, groups, eps, attn):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(in_channels=ch, **attn)])


class UpBlock(_Container):
    def __init__(self, in_ch, out_ch, prev_ch, temb_ch, num_layers, add_upsample, groups, eps, attn=None):
        super().__init__()
        res = []
        for i in range(num_layers):
            skip = in_ch if i == num_layers - 1 else out_ch
            rin = prev_ch if i == 0 else out_ch
            res.append(ResnetBlock2D(rin + skip, out_ch, temb_ch, groups, eps))
        self.resnets = nn.ModuleList(res)
        self.attentions = (nn.ModuleList([Transformer2DModel(in_channels=out_ch, **attn) for _ in range(num_layers)])
                           if attn is not None else None)
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None


class DiffusersUNet2DCondWrapper(nn.Module):
    """Same constructor kwargs, state-dict keys and `forward` contract as the reference wrapper
    (src/flash/models/unets/unet.py:62-127)."""

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
        if mid_block_type != "UNetMidBlock2DCrossAttn" or act_fn != "silu" or only_cross_attention:
            raise NotImplementedError("only the UNet variants used by the Flash-Diffusion examples are built")
        if not flip_sin_to_cos or freq_shift != 0:
            raise NotImplementedError("timestep embedding kernel implements flip_sin_to_cos=True, freq_shift=0")
        n = len(down_block_types)
        boc = list(block_out_channels)
        heads = _per_block(num_attention_heads or attention_head_dim, n)     # diffusers naming quirk
        tlpb = _per_block(transformer_layers_per_block, n)
        lpb = _per_block(layers_per_block, n)
        temb_ch = boc[0] * 4
        self.in_channels, self.out_channels = in_channels, out_channels
        self.time_dim = boc[0]
        self.center_input_sample = center_input_sample
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        if class_embed_type == "projection":
            self.class_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, temb_ch)
        elif class_embed_type is None:
            self.class_embedding = None
        else:
            raise NotImplementedError(f"class_embed_type={class_embed_type}")

        def attn_kwargs(h, ch, layers):
            return dict(heads=h, dim_head=ch // h, num_layers=layers, cross_attention_dim=cross_attention_dim,
                        groups=norm_num_groups, use_linear_projection=use_linear_projection)

        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            if t not in ("CrossAttnDownBlock2D", "DownBlock2D"):
                raise NotImplementedError(t)
            attn = attn_kwargs(heads[i], out_ch, tlpb[i]) if t == "CrossAttnDownBlock2D" else None
            self.down_blocks.append(DownBlock(in_ch, out_ch, temb_ch, lpb[i], i != n - 1, norm_num_groups, norm_eps, attn))
        self.mid_block = MidBlock(boc[-1], temb_ch, norm_num_groups, norm_eps, attn_kwargs(heads[-1], boc[-1], tlpb[-1]))
        self.up_blocks = nn.ModuleList()
        rboc, rheads, rtl, rlpb = boc[::-1], heads[::-1], tlpb[::-1], lpb[::-1]
        if reverse_transformer_layers_per_block is not None:
            rtl = _per_block(reverse_transformer_layers_per_block, n)
        out_ch = rboc[0]
        for i, t in enumerate(up_block_types):
            prev_ch, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, n - 1)]
            if t not in ("CrossAttnUpBlock2D", "UpBlock2D"):
                raise NotImplementedError(t)
            attn = attn_kwargs(rheads[i], out_ch, rtl[i]) if t == "CrossAttnUpBlock2D" else None
            self.up_blocks.append(UpBlock(in_ch, out_ch, prev_ch, temb_ch, rlpb[i] + 1, i != n - 1,
                                          norm_num_groups, norm_eps, attn))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.__dict__["_packs"] = {}

    # ------------------------------------------------------------------------------------ helpers
    def _pack(self, key, make):
        packs = self.__dict__.setdefault("_packs", {})
        if key not in packs:
            packs[key] = make()
        return packs[key]

    def __deepcopy__(self, memo):
        # kernel-side pack caches hold device buffers tied to the source parameters: never copy them
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_packs", "_kv_store"):
                new.__dict__[k] = {}
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        for m in new.modules():
            m.__dict__.pop("_fd_cache", None)
        return new

    def freeze(self):
        self.eval()
        for p in self.parameters():
            p.requires_grad = False

    def add_adapter(self, lora_config):
        """diffusers `add_adapter` (peft inject_adapter_in_model) — reference call examples/train_flash_sdxl.py:217."""
        inject_lora(self, lora_config)
        # The UNet engine executes LoRA on the attention projections, proj_in/proj_out and ff.net.2 (ops.linear).  The
        # GEGLU projection and the time / class embedding MLPs run fused, LoRA-free GEMMs: a target list reaching them
        # would create adapters that are never applied and never receive gradients — refuse instead.
        bad = [n for n, m in self.named_modules() if isinstance(m, LoRALinear) and
               (n.endswith("ff.net.0.proj") or n.endswith("time_emb_proj") or n.startswith("time_embedding.")
                or n.startswith("class_embedding."))]
        if bad:
            raise NotImplementedError(f"LoRA on {bad[:3]}... is not executed by the B200 UNet engine (GEGLU / "
                                      "time-embedding GEMMs are fused without an adapter segment); the reference's "
                                      "SD / SDXL target lists (to_q, to_k, to_v, to_out.0) do not reach them")
        self.__dict__["_packs"] = {}
        return self

    # ------------------------------------------------------------------------------------ engine
    def _temb(self, timestep, class_labels, B, device):
        """emb = time_embedding(sinusoid(t)) [+ class_embedding(vector)]  -> SiLU(emb) bf16 [B, temb] (every consumer
        applies SiLU first: ResnetBlock2D.time_emb_proj(silu(emb)))."""
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([float(timestep)], dtype=torch.float32, device=device)
        timestep = timestep.to(device=device, dtype=torch.float32).reshape(-1)
        if timestep.numel() == 1 and B > 1:
            timestep = timestep.expand(B)
        timestep = timestep.contiguous()
        te = self.time_embedding
        t_emb = raw.timestep_embedding(timestep, self.time_dim)
        l1 = self._pack("te1", lambda: LinearPack(te.linear_1))
        l2 = self._pack("te2", lambda: LinearPack(te.linear_2))
        p1, p2 = l1.pack(), l2.pack()
        h = raw.silu_f32_to_bf16(raw.gemm(t_emb, p1["w"], bias=p1["b"], out_fp32=True))
        if self.class_embedding is None:
            emb = raw.gemm(h, p2["w"], bias=p2["b"], out_fp32=True)
        else:
            if class_labels is None:
                raise ValueError("class_labels should be provided when num_class_embeds > 0")
            ce = self.class_embedding
            c1 = self._pack("ce1", lambda: LinearPack(ce.linear_1))
            c2 = self._pack("ce2", lambda: LinearPack(ce.linear_2))
            q1, q2 = c1.pack(), c2.pack()
            v = c1.pad_k(raw.cast_scale(class_labels.detach().float().contiguous(), 1.0))   # odd widths: zero columns
            hc = raw.silu_f32_to_bf16(raw.gemm(v, q1["w"], bias=q1["b"], out_fp32=True))
            # emb = time_embedding(t) + class_embedding(vector): both second Linears in ONE GEMM (two K segments)
            bsum = self._pack("emb_bias", lambda: (p2["b"] + q2["b"]).contiguous())
            emb = raw.gemm(h, p2["w"], a2=hc, b2=q2["w"], bias=bsum, out_fp32=True)
        return raw.silu_f32_to_bf16(emb)

    def _all_resnets(self):
        out = []
        for blk in list(self.down_blocks) + [self.mid_block] + list(self.up_blocks):
            out.extend(blk.resnets)
        return out

    def _temb_rows(self, emb_silu):
        """All `time_emb_proj` of the network as ONE GEMM: [B, temb] x [sum(C_out), temb]^T -> per-resnet views."""
        resnets = self._all_resnets()
        pack = self._pack("temb_all", lambda: LinearPack([r.time_emb_proj for r in resnets]))
        p = pack.pack()
        rows = raw.gemm(emb_silu, p["w"], bias=p["b"], out_fp32=True)
        out, off = {}, 0
        for r in resnets:
            c = r.time_emb_proj.weight.shape[0]
            out[id(r)] = rows[:, off:off + c]
            off += c
        return out

    def _resnet(self, r, x, geom, trow):
        c1 = self._pack(("c1", id(r)), lambda: ConvPack(r.conv1))
        c2 = self._pack(("c2", id(r)), lambda: ConvPack(r.conv2, r.conv_shortcut))
        # `arena`: the conv epilogues leave the per-image column sums of their outputs, so the GroupNorm that follows
        # (norm2 here, norm1 / Transformer2DModel.norm / conv_norm_out of the next block) needs no reduction pass
        arena = self.__dict__.get("_arena")
        h = ops.group_norm(x, geom, r.norm1, silu=True)
        h = ops.conv3x3(h, geom, c1, rowvec=trow, arena=arena)
        h = ops.group_norm(h, geom, r.norm2, silu=True)
        if r.conv_shortcut is not None:
            return ops.conv3x3(h, geom, c2, x2=x, arena=arena)
        return ops.conv3x3(h, geom, c2, residual=x, arena=arena)

    def _attention(self, a, x, ctx, B, residual, norm, stats):
        """x: the residual stream (un-normalised), `norm` its LayerNorm.  When the producer GEMM left row statistics
        (`stats`) and no gradient / LoRA is involved, the LayerNorm is folded into the projection GEMM; otherwise the
        LayerNorm kernel runs.  Returns (new residual stream, its row statistics or None)."""
        H, d = a.heads, a.dim_head
        dp = (d + 15) // 16 * 16                       # head dim the attention kernel sees (zero-padded channels)
        hp = (H, d, dp) if dp != d else None
        scale = d ** -0.5
        inner = H * dp
        first = self._pack(("qkv" if not a.is_cross else "q", id(a)),
                           lambda: LinearPack([a.to_q, a.to_k, a.to_v] if not a.is_cross else a.to_q, head_pad=hp))
        if ops.ln_foldable(x, stats, first):
            proj = ops.linear_ln(x, stats, norm, first)
        else:
            proj = ops.linear(ops.layer_norm(x, norm), first)
        if not a.is_cross:
            o = ops.attention_self(proj.view(B, -1, 3 * inner), H, head_dim=dp, scale=scale).view(-1, inner)
        else:
            kvp = self._pack(("kv", id(a)), lambda: LinearPack([a.to_k, a.to_v], head_pad=hp))
            mode = self.__dict__.get("_kv_mode")
            if mode is not None and not torch.is_grad_enabled():
                # cross-attention K/V depend on the text conditioning only: computed by the first evaluation of a
                # frozen-teacher rollout ("fill") into persistent buffers and re-read by the following ones ("reuse")
                store = self.__dict__.setdefault("_kv_store", {})
                key = (id(a), ctx.shape[0])
                if mode == "reuse":
                    kv = store[key]
                else:
                    buf = store.get(key)
                    if buf is None:
                        buf = store[key] = torch.empty((ctx.shape[0], 2 * inner), device=ctx.device,
                                                       dtype=torch.bfloat16)
                    kv = ops._linear_fwd_raw(kvp.pad_k(ctx), kvp, None, out=buf)[0]
            else:
                kv = ops.linear(ctx, kvp)
            o = ops.attention_cross(proj.view(B, -1, inner), kv.view(B, -1, 2 * inner), H, head_dim=dp,
                                    scale=scale).view(-1, inner)
        return ops.linear(o, self._pack(("o", id(a)), lambda: LinearPack(a.to_out[0], head_pad=hp, pad_cols=True)),
                          residual=residual, want_stats=True, arena=self.__dict__.get("_arena"))

    def _transformer(self, t, x, geom, ctx):
        B = geom[0]
        h = ops.group_norm(x, geom, t.norm, silu=False)
        arena = self.__dict__.get("_arena")
        h, st = ops.linear(h, self._pack(("pi", id(t)), lambda: LinearPack(t.proj_in)), want_stats=True, arena=arena)
        n = len(t.transformer_blocks)
        for i, blk in enumerate(t.transformer_blocks):
            h, st = self._attention(blk.attn1, h, None, B, h, blk.norm1, st)
            h, st = self._attention(blk.attn2, h, ctx, B, h, blk.norm2, st)
            ff1 = self._pack(("ff1", id(blk)), lambda: LinearPack(blk.ff.net[0].proj, geglu=True))
            if ops.ln_foldable(h, st, ff1):
                g = ops.linear_ln(h, st, blk.norm3, ff1)
            else:
                g = ops.geglu(ops.layer_norm(h, blk.norm3), ff1)
            h, st = ops.linear(g, self._pack(("ff2", id(blk)), lambda: LinearPack(blk.ff.net[2])), residual=h,
                               want_stats=(i + 1 < n), arena=arena)
        return ops.linear(h, self._pack(("po", id(t)), lambda: LinearPack(t.proj_out)), residual=x, arena=arena,
                          colstats_images=B)

    supports_kv_cache = True

    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                conditioning: Dict[str, torch.Tensor], down_intrablock_additional_residuals=None,
                return_intermediate: bool = False, *args, kv_cache=None, **kwargs):
        """`kv_cache` (B200 extension, no-grad evaluations only): "fill" stores the cross-attention K/V projections of
        this call's text conditioning, "reuse" reads them instead of recomputing the 70 K/V GEMMs — for consecutive
        evaluations with THE SAME conditioning (the teacher's CFG rollout, flash_diffusion_model.py:288-324)."""
        if kv_cache not in (None, "fill", "reuse"):
            raise ValueError(f"kv_cache={kv_cache!r}")
        self.__dict__["_kv_mode"] = kv_cache if not torch.is_grad_enabled() else None
        self.__dict__["_arena"] = ops.StatsArena(sample.device) if sample.is_cuda else None
        try:
            return self._forward(sample, timestep, conditioning, down_intrablock_additional_residuals,
                                 return_intermediate)
        finally:
            self.__dict__["_kv_mode"] = None
            self.__dict__["_arena"] = None

    def _forward(self, sample, timestep, conditioning, down_intrablock_additional_residuals=None,
                 return_intermediate=False):
        assert isinstance(conditioning, dict), "conditionings must be a dictionary"
        if down_intrablock_additional_residuals is not None:
            raise NotImplementedError("T2I-adapter residuals are out of scope of the B200 hot path (SURVEY §2 row 7)")
        if not sample.is_cuda:
            raise RuntimeError("DiffusersUNet2DCondWrapper runs only on CUDA (B200) tensors: there is no CPU fallback")
        cond = conditioning["cond"]
        class_labels, crossattn, concat = cond.get("vector"), cond.get("crossattn"), cond.get("concat")
        if concat is not None:
            sample = torch.cat([sample, concat], dim=1)
        if self.center_input_sample:
            sample = 2 * sample - 1.0
        NB, Cin, H, W = sample.shape
        dev = sample.device
        with torch.no_grad():
            trows = self._temb_rows(self._temb(timestep, class_labels, NB, dev))
            ctx = None
            if crossattn is not None:
                ctx = raw.cast_scale(crossattn.detach().float().contiguous().view(-1, crossattn.shape[-1]), 1.0)

        conv_in = self._pack("conv_in", lambda: ConvPack(self.conv_in))
        geom = (NB, H, W)
        arena = self.__dict__.get("_arena")
        x = ops.conv3x3(ops.to_nhwc(sample, conv_in.cin), geom, conv_in, arena=arena)
        skips = [(x, geom)]
        for blk in self.down_blocks:
            for i, r in enumerate(blk.resnets):
                x = self._resnet(r, x, geom, trows[id(r)])
                if blk.attentions is not None:
                    x = self._transformer(blk.attentions[i], x, geom, ctx)
                skips.append((x, geom))
            if blk.downsamplers is not None:
                ds = blk.downsamplers[0]
                x = ops.conv3x3(x, geom, self._pack(("ds", id(ds)), lambda: ConvPack(ds.conv)), stride=2, arena=arena)
                geom = (NB, geom[1] // 2, geom[2] // 2)
                skips.append((x, geom))
        mb = self.mid_block
        x = self._resnet(mb.resnets[0], x, geom, trows[id(mb.resnets[0])])
        x = self._transformer(mb.attentions[0], x, geom, ctx)
        x = self._resnet(mb.resnets[1], x, geom, trows[id(mb.resnets[1])])
        if return_intermediate:
            # fork-only kwarg (reference unet.py:72,116): mid-block features, NCHW
            return ops.to_nchw(x, geom, x.shape[1])
        for blk in self.up_blocks:
            for i, r in enumerate(blk.resnets):
                s, sgeom = skips.pop()
                assert sgeom == geom
                x = self._resnet(r, ops.concat(x, s), geom, trows[id(r)])
                if blk.attentions is not None:
                    x = self._transformer(blk.attentions[i], x, geom, ctx)
            if blk.upsamplers is not None:
                us = blk.upsamplers[0]
                x = ops.upsample2x(x, geom)
                geom = (NB, geom[1] * 2, geom[2] * 2)
                x = ops.conv3x3(x, geom, self._pack(("us", id(us)), lambda: ConvPack(us.conv)), arena=arena)
        x = ops.group_norm(x, geom, self.conv_norm_out, silu=True)
        conv_out = self._pack("conv_out", lambda: ConvPack(self.conv_out))
        y = ops.conv3x3(x, geom, conv_out, out_fp32=True)
        return ops.to_nchw(y, geom, self.out_channels)
