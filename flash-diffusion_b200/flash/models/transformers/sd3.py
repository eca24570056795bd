"""B200-native `DiffusersSD3Transformer2DWrapper` — SD3 MMDiT (reference src/flash/models/transformers/tranformers.py:103-163;
constructor kwargs as at examples/train_flash_sd3.py:65-77).  Forward and backward: activation gradients (incl. the
input gradient the GAN generator turn needs through the frozen backbone) and LoRA gradients for every Linear the
reference's target list names (attention, feed-forward, AdaLN and embedding linears) and for the 2x2 patch convolution,
which the "proj" target also names (peft `lora.Conv2d`).

Kernel mapping (UPSTREAM diffusers `SD3Transformer2DModel` math, restated in oracle/sd3.py):
  PatchEmbed conv 2x2/2 + cropped sin-cos table   space-to-depth + 4-tap implicit GEMM, table added as the epilogue residual
  timestep + pooled-text embedding                 small fd_gemm launches (M = batch)
  every AdaLN-Zero / AdaLN-continuous `linear`     ONE fd_gemm over the concatenated weights of all 24 blocks
  LN * (1+scale) + shift (image and text streams)  fd_layernorm_modulate
  joint attention (24 heads x 64)                  fused q|k|v GEMMs per stream, token concat, fd_attn_fwd (tuned d=64 kernel)
  gate * f(x) + x                                  fd_gemm epilogue (bias -> gelu-tanh -> per-sample gate -> residual);
                                                   with gradients: fd_gemm, then fd_gate_residual (keeps f(x) for d gate)
  un-patchify                                      fd_unpatchify / fd_patchify
  backward                                         fd_layernorm_modulate_bwd, fd_gate_bwd, fd_gelu_tanh_bwd (pre-activation
                                                   recomputed), fd_attn_bwd, fd_gemm for every dX / LoRA dA, dB
"""
from typing import Dict, Optional, Union

import torch
import torch.nn as nn

from ...b200 import ops, raw
from ...b200.ops import LinearPack, cache_of
from ..lora import inject_lora
from ..unets.unet import TimestepEmbedding, _Container
from .transformers import FeedForward, patch_embed, patch_lora_pack, sincos_2d


class PatchEmbedSD3(_Container):
    def __init__(self, sample_size, patch_size, in_channels, embed_dim, pos_embed_max_size):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, patch_size, stride=patch_size)
        self.max = pos_embed_max_size
        pe = sincos_2d(embed_dim, pos_embed_max_size, base_size=sample_size // patch_size, interpolation_scale=1)
        self.register_buffer("pos_embed", torch.from_numpy(pe).float()[None], persistent=True)

    def cropped(self, h, w):
        top, left = (self.max - h) // 2, (self.max - w) // 2
        return self.pos_embed.reshape(self.max, self.max, -1)[top:top + h, left:left + w].reshape(h * w, -1)


class TextProjSilu(_Container):
    def __init__(self, in_features, hidden):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden)
        self.linear_2 = nn.Linear(hidden, hidden)


class CombinedTimestepTextProjEmbeddings(_Container):
    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, embedding_dim)
        self.text_embedder = TextProjSilu(pooled_projection_dim, embedding_dim)


class AdaLinear(_Container):
    def __init__(self, dim, chunks):
        super().__init__()
        self.linear = nn.Linear(dim, chunks * dim)


class JointAttention(_Container):
    def __init__(self, dim, heads, dim_head, context_pre_only):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        self.add_k_proj, self.add_v_proj, self.add_q_proj = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])
        self.to_add_out = None if context_pre_only else nn.Linear(inner, dim)


class JointBlock(_Container):
    def __init__(self, dim, heads, dim_head, context_pre_only):
        super().__init__()
        self.pre_only = context_pre_only
        self.norm1 = AdaLinear(dim, 6)
        self.norm1_context = AdaLinear(dim, 2 if context_pre_only else 6)
        self.attn = JointAttention(dim, heads, dim_head, context_pre_only)
        self.ff = FeedForward(dim)
        self.ff_context = None if context_pre_only else FeedForward(dim)


class DiffusersSD3Transformer2DWrapper(nn.Module):
    def __init__(self, sample_size=128, patch_size=2, in_channels=16, num_layers=18, attention_head_dim=64,
                 num_attention_heads=18, joint_attention_dim=4096, caption_projection_dim=1152,
                 pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=96, **unused):
        super().__init__()
        if attention_head_dim != 64 or patch_size != 2:
            raise NotImplementedError("SD3 MMDiT is built for head dim 64, patch size 2")
        D = num_attention_heads * attention_head_dim
        if caption_projection_dim != D:
            raise NotImplementedError("caption_projection_dim must equal the inner dim")
        self.inner_dim, self.patch_size, self.out_channels, self.in_channels = D, patch_size, out_channels, in_channels
        self.heads = num_attention_heads
        self.pos_embed = PatchEmbedSD3(sample_size, patch_size, in_channels, D, pos_embed_max_size)
        self.time_text_embed = CombinedTimestepTextProjEmbeddings(D, pooled_projection_dim)
        self.context_embedder = nn.Linear(joint_attention_dim, caption_projection_dim)
        self.transformer_blocks = nn.ModuleList(
            [JointBlock(D, num_attention_heads, attention_head_dim, context_pre_only=(i == num_layers - 1))
             for i in range(num_layers)])
        self.norm_out = AdaLinear(D, 2)
        self.proj_out = nn.Linear(D, patch_size * patch_size * out_channels)
        self.__dict__["_packs"] = {}

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
        inject_lora(self, lora_config)
        self.__dict__["_packs"] = {}
        return self

    @staticmethod
    def _lin(x, pack: LinearPack, **epi):
        p = pack.pack()
        a2 = b2 = None
        if pack.has_lora:
            lp = pack.pack_lora()
            a2, b2 = raw.gemm(x, lp["a"]), lp["b"]
        return raw.gemm(x, p["w"], a2=a2, b2=b2, bias=p["b"], **epi)

    def _all_ada(self):
        mods = []
        for blk in self.transformer_blocks:
            mods += [blk.norm1, blk.norm1_context]
        return mods + [self.norm_out]

    @staticmethod
    def _silu_bf16(x):
        """silu(fp32) -> bf16 on a [B, D] embedding; with a gradient it is three torch ops on B rows."""
        if torch.is_grad_enabled() and x.requires_grad:
            return torch.nn.functional.silu(x).to(torch.bfloat16)
        return raw.silu_f32_to_bf16(x)

    def _patch_embed(self, sample, B, Cin, H, W):
        pe, D, dev = self.pos_embed, self.inner_dim, sample.device
        hh, ww = H // 2, W // 2
        N = hh * ww
        cpad = (Cin + 7) // 8 * 8

        def build_patch():
            wt = pe.proj.weight.detach().float()
            buf = torch.zeros((D, 4, 64), device=dev)
            buf[:, :, :Cin] = wt.permute(0, 2, 3, 1).reshape(D, 4, Cin)
            w = raw.cast_scale(buf.reshape(D, 256), 1.0)
            return {"w": w, "w_t": raw.cast_scale(buf.reshape(D, 256).t().contiguous(), 1.0),
                    "b": pe.proj.bias.detach().float().contiguous(),
                    "pos": raw.cast_scale(pe.cropped(hh, ww).to(dev).contiguous(), 1.0)}
        pk = cache_of(pe.proj).get(("patch", hh, ww), [pe.proj.weight, pe.proj.bias], build_patch)
        pos_b = self._pack(("pos_tiled", B, N), lambda: pk["pos"].repeat(B, 1).contiguous())
        return patch_embed(sample.float(), pk, pos_b, (B, Cin, H, W, cpad), patch_lora_pack(pe.proj, Cin, dev))

    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                conditioning: Dict[str, torch.Tensor], hidden_states_masks: Optional[torch.Tensor] = None,
                *args, **kwargs):
        assert isinstance(conditioning, dict), "conditionings must be a dictionary"
        if not sample.is_cuda:
            raise RuntimeError("DiffusersSD3Transformer2DWrapper runs only on CUDA (B200) tensors: there is no CPU fallback")
        cond = conditioning["cond"]
        pooled, crossattn, concat = cond.get("vector"), cond.get("crossattn"), cond.get("concat")
        c_keep = sample.shape[1]
        if concat is not None:
            sample = torch.cat([sample, concat], dim=1)
        B, Cin, H, W = sample.shape
        dev, p, D, Hh = sample.device, self.patch_size, self.inner_dim, self.heads
        hh, ww = H // p, W // p
        N, T = hh * ww, crossattn.shape[1]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([float(timestep)], dtype=torch.float32, device=dev)
        timestep = timestep.detach().to(device=dev, dtype=torch.float32).reshape(-1)
        if timestep.numel() == 1 and B > 1:
            timestep = timestep.expand(B)
        timestep = timestep.contiguous()
        tt = self.time_text_embed
        te1 = self._pack("te1", lambda: LinearPack(tt.timestep_embedder.linear_1))
        te2 = self._pack("te2", lambda: LinearPack(tt.timestep_embedder.linear_2))
        tx1 = self._pack("tx1", lambda: LinearPack(tt.text_embedder.linear_1))
        tx2 = self._pack("tx2", lambda: LinearPack(tt.text_embedder.linear_2))
        ada = self._all_ada()
        ada_pack = self._pack("ada_all", lambda: LinearPack([m.linear for m in ada]))
        # `tuning` = the embedding / AdaLN linears carry trainable adapters (the reference's SD3 LoRA targets do): their
        # gradients need per-module products, so the one-GEMM fusions below are used only without them
        tuning = torch.is_grad_enabled() and any(q.requires_grad for pk in (te1, te2, tx1, tx2, ada_pack)
                                                 for q in pk.lora_params())
        t_in = raw.timestep_embedding(timestep, 256)
        p_in = raw.cast_scale(pooled.detach().float().contiguous(), 1.0)
        ht = self._silu_bf16(ops.linear(t_in, te1, out_fp32=True))
        hp_ = self._silu_bf16(ops.linear(p_in, tx1, out_fp32=True))
        if tuning:
            temb = ops.linear(ht, te2, out_fp32=True) + ops.linear(hp_, tx2, out_fp32=True)
            s_temb = self._silu_bf16(temb)
            mods = [ops.linear(s_temb, self._pack(("ada", id(m)), lambda m=m: LinearPack(m.linear)), out_fp32=True)
                    .view(B, -1, D) for m in ada]
            mod = lambda idx: mods[idx]
        else:
            # temb = timestep_embedder(sinusoid(t)) + text_embedder(pooled): both second Linears as ONE two-segment
            # GEMM, then every AdaLN linear of the network in one GEMM
            p2, q2 = te2.pack(), tx2.pack()
            bsum = self._pack("temb_bias", lambda: (p2["b"] + q2["b"]).contiguous())
            if te2.has_lora or tx2.has_lora:
                temb = self._lin(ht, te2, out_fp32=True) + self._lin(hp_, tx2, out_fp32=True)
            else:
                temb = raw.gemm(ht, p2["w"], a2=hp_, b2=q2["w"], bias=bsum, out_fp32=True)
            mod_all = self._lin(raw.silu_f32_to_bf16(temb), ada_pack, out_fp32=True)          # [B, sum(chunks)*D]
            offs, off = [], 0
            for m in ada:
                n_out = m.linear.weight.shape[0]
                offs.append((off, n_out // D))
                off += n_out

            def mod(idx):
                o, k = offs[idx]
                return mod_all[:, o:o + k * D].view(B, k, D)

        c = ops.linear(raw.cast_scale(crossattn.detach().float().contiguous().view(B * T, -1), 1.0),
                       self._pack("ctx", lambda: LinearPack(self.context_embedder)))
        h = self._patch_embed(sample, B, Cin, H, W)

        inner = Hh * 64
        for li, blk in enumerate(self.transformer_blocks):
            mx, mc = mod(2 * li), mod(2 * li + 1)        # image: shift,scale,gate (msa), shift,scale,gate (mlp)
            a = blk.attn
            nx = ops.modulate(h, mx[:, 1], mx[:, 0], N)
            if blk.pre_only:
                nc = ops.modulate(c, mc[:, 0], mc[:, 1], T)           # AdaLN-continuous: (scale, shift)
            else:
                nc = ops.modulate(c, mc[:, 1], mc[:, 0], T)
            qkv_x = ops.linear(nx, self._pack(("qkv", id(a)), lambda: LinearPack([a.to_q, a.to_k, a.to_v])))
            qkv_c = ops.linear(nc, self._pack(("aqkv", id(a)), lambda: LinearPack([a.add_q_proj, a.add_k_proj, a.add_v_proj])))
            joint = torch.cat([qkv_x.view(B, N, 3 * inner), qkv_c.view(B, T, 3 * inner)], dim=1)
            o = ops.attention_self(joint, Hh)
            ox = o[:, :N].reshape(B * N, inner)
            h = ops.gated_linear(ox, self._pack(("o", id(a)), lambda: LinearPack(a.to_out[0])), mx[:, 2], h, N)
            n2 = ops.modulate(h, mx[:, 4], mx[:, 3], N)
            f = ops.linear(n2, self._pack(("ff1", id(blk)), lambda: LinearPack(blk.ff.net[0].proj)), act=1)
            h = ops.gated_linear(f, self._pack(("ff2", id(blk)), lambda: LinearPack(blk.ff.net[2])), mx[:, 5], h, N)
            if not blk.pre_only:
                oc = o[:, N:].reshape(B * T, inner)
                c = ops.gated_linear(oc, self._pack(("ao", id(a)), lambda: LinearPack(a.to_add_out)), mc[:, 2], c, T)
                nc2 = ops.modulate(c, mc[:, 4], mc[:, 3], T)
                fc = ops.linear(nc2, self._pack(("cff1", id(blk)), lambda: LinearPack(blk.ff_context.net[0].proj)), act=1)
                c = ops.gated_linear(fc, self._pack(("cff2", id(blk)), lambda: LinearPack(blk.ff_context.net[2])),
                                     mc[:, 5], c, T)
        mo = mod(len(ada) - 1)                                                          # (scale, shift)
        nf = ops.modulate(h, mo[:, 0], mo[:, 1], N)
        out = ops.linear(nf, self._pack("proj_out", lambda: LinearPack(self.proj_out)), out_fp32=True)
        return ops.unpatchify(out, B, hh, ww, p, self.out_channels, c_keep)
