"""GPU parity of the hand-written backward kernels against torch autograd (fp32 math)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _cos(a, b):
    return F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


@pytest.mark.parametrize("B,H,Nq,Nkv", [(1, 1, 128, 128), (2, 3, 256, 384), (1, 5, 1024, 1024), (2, 2, 512, 77),
                                        (1, 2, 200, 333)])
def test_attention_bwd(B, H, Nq, Nkv):
    from flash.b200 import raw
    torch.manual_seed(B + H + Nq + Nkv)
    q = torch.randn(B, Nq, H * 64, device="cuda").bfloat16()
    k = torch.randn(B, Nkv, H * 64, device="cuda").bfloat16()
    v = torch.randn(B, Nkv, H * 64, device="cuda").bfloat16()
    do = torch.randn(B, Nq, H * 64, device="cuda").bfloat16()
    o, lse = raw.attention_fwd(q, k, v, H, need_lse=True)
    dq, dk, dv = raw.attention_bwd(q, k, v, o, lse, do, H)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf.view(B, Nq, H, 64).transpose(1, 2), kf.view(B, Nkv, H, 64).transpose(1, 2),
                                         vf.view(B, Nkv, H, 64).transpose(1, 2)).transpose(1, 2).reshape(B, Nq, H * 64)
    ref.backward(do.float())
    assert _rel(dq, qf.grad) < 2e-2, _rel(dq, qf.grad)
    assert _rel(dk, kf.grad) < 2e-2, _rel(dk, kf.grad)
    assert _rel(dv, vf.grad) < 2e-2, _rel(dv, vf.grad)


def _sdpa_ref(q, k, v, do, H, d, scale, kv_len):
    B, Nq, Nkv = q.shape[0], q.shape[1], k.shape[1]
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    mask = None
    if kv_len is not None:
        mask = (torch.arange(Nkv, device="cuda")[None, :] < kv_len[:, None].long())[:, None, None, :]
    ref = F.scaled_dot_product_attention(qf.view(B, Nq, H, d).transpose(1, 2), kf.view(B, Nkv, H, d).transpose(1, 2),
                                         vf.view(B, Nkv, H, d).transpose(1, 2), attn_mask=mask, scale=scale)
    ref = ref.transpose(1, 2).reshape(B, Nq, H * d)
    ref.backward(do.float())
    return ref, qf.grad, kf.grad, vf.grad


@pytest.mark.parametrize("B,H,d,Nq,Nkv,masked", [
    (2, 2, 48, 256, 256, False),      # SD1.5 level 1 (40 -> 48), one sub-tile
    (2, 3, 80, 200, 333, False),      # SD1.5 level 2 / PixArt (72 -> 80): 64 + 16 columns
    (2, 2, 80, 1024, 1024, False),
    (2, 4, 80, 256, 120, True),       # PixArt cross-attention with a T5 padding mask
    (2, 2, 64, 256, 120, True),       # d = 64 with a mask
    (2, 2, 16, 128, 77, False),
    (2, 8, 160, 256, 256, False),     # SD1.5 level 3 self-attention (16x16 tokens): CUDA-core passes
    (2, 8, 160, 64, 77, True),        # SD1.5 mid block cross-attention
])
def test_attention_bwd_generic(B, H, d, Nq, Nkv, masked):
    from flash.b200 import raw
    torch.manual_seed(B + H + d + Nq + Nkv)
    q = torch.randn(B, Nq, H * d, device="cuda").bfloat16()
    k = torch.randn(B, Nkv, H * d, device="cuda").bfloat16()
    v = torch.randn(B, Nkv, H * d, device="cuda").bfloat16()
    do = torch.randn(B, Nq, H * d, device="cuda").bfloat16()
    kv_len = torch.tensor([Nkv - 43, Nkv][:B], device="cuda", dtype=torch.int32) if masked else None
    scale = 0.9 * d ** -0.5
    o, lse = raw.attention_fwd(q, k, v, H, scale=scale, need_lse=True, head_dim=d, kv_len=kv_len)
    dq, dk, dv = raw.attention_bwd(q, k, v, o, lse, do, H, scale=scale, head_dim=d, kv_len=kv_len)
    ref, gq, gk, gv = _sdpa_ref(q, k, v, do, H, d, scale, kv_len)
    assert _rel(o, ref) < 2e-2
    assert _rel(dq, gq) < 2e-2, _rel(dq, gq)
    assert _rel(dk, gk) < 2e-2, _rel(dk, gk)
    assert _rel(dv, gv) < 2e-2, _rel(dv, gv)
    if masked:      # padded keys get exactly zero gradient
        assert float(dk[0, Nkv - 43:].abs().max()) == 0 and float(dv[0, Nkv - 43:].abs().max()) == 0


def test_attention_generic_fused_autograd_views():
    """Strided q|k|v views of one fused projection buffer, as the UNet / DiT engines pass them."""
    from flash.b200 import ops
    torch.manual_seed(0)
    B, N, H, d = 2, 200, 3, 80
    qkv = torch.randn(B, N, 3 * H * d, device="cuda").bfloat16().requires_grad_(True)
    o = ops.attention_self(qkv, H, head_dim=d, scale=72 ** -0.5)
    g = torch.randn_like(o)
    o.backward(g)
    qf = qkv.detach().float().requires_grad_(True)
    q, k, v = qf.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    F.scaled_dot_product_attention(q, k, v, scale=72 ** -0.5).transpose(1, 2).reshape(B, N, H * d).backward(g.float())
    assert _rel(qkv.grad, qf.grad) < 2e-2
    q2 = torch.randn(B, N, H * d, device="cuda").bfloat16().requires_grad_(True)
    kv = torch.randn(B, 120, 2 * H * d, device="cuda").bfloat16().requires_grad_(True)
    kv_len = torch.tensor([77, 120], device="cuda", dtype=torch.int32)
    ops.attention_cross(q2, kv, H, head_dim=d, kv_len=kv_len).backward(g)
    qf2, kvf = q2.detach().float().requires_grad_(True), kv.detach().float().requires_grad_(True)
    k2, v2 = kvf.view(B, 120, 2, H, d).permute(2, 0, 3, 1, 4)
    mask = (torch.arange(120, device="cuda")[None, :] < kv_len[:, None].long())[:, None, None, :]
    F.scaled_dot_product_attention(qf2.view(B, N, H, d).transpose(1, 2), k2, v2, attn_mask=mask) \
        .transpose(1, 2).reshape(B, N, H * d).backward(g.float())
    assert _rel(q2.grad, qf2.grad) < 2e-2 and _rel(kv.grad, kvf.grad) < 2e-2


def test_attention_fused_autograd():
    from flash.b200 import ops
    torch.manual_seed(0)
    B, N, H = 2, 256, 2
    qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16().requires_grad_(True)
    o = ops.attention_self(qkv, H)
    g = torch.randn_like(o)
    o.backward(g)
    qf = qkv.detach().float().requires_grad_(True)
    q, k, v = qf.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, H * 64)
    ref.backward(g.float())
    assert _rel(qkv.grad, qf.grad) < 2e-2
    q2 = torch.randn(B, N, H * 64, device="cuda").bfloat16().requires_grad_(True)
    kv = torch.randn(B, 77, 2 * H * 64, device="cuda").bfloat16().requires_grad_(True)
    o2 = ops.attention_cross(q2, kv, H)
    o2.backward(g)
    qf2, kvf = q2.detach().float().requires_grad_(True), kv.detach().float().requires_grad_(True)
    k2, v2 = kvf.view(B, 77, 2, H, 64).permute(2, 0, 3, 1, 4)
    ref2 = F.scaled_dot_product_attention(qf2.view(B, N, H, 64).transpose(1, 2), k2, v2).transpose(1, 2).reshape(B, N, H * 64)
    ref2.backward(g.float())
    assert _rel(q2.grad, qf2.grad) < 2e-2 and _rel(kv.grad, kvf.grad) < 2e-2


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(2, 32, 32, 64, 128), (1, 64, 64, 320, 320), (2, 16, 16, 192, 64)])
def test_conv_dgrad(stride, NB, H, W, Cin, Cout):
    from flash.b200 import ops
    torch.manual_seed(stride + Cin)
    torch.backends.cudnn.allow_tf32 = False
    conv = torch.nn.Conv2d(Cin, Cout, 3, stride=stride, padding=1).cuda()
    pack = ops.ConvPack(conv)
    x = torch.randn(NB, Cin, H, W, device="cuda").bfloat16()
    x_nhwc = x.permute(0, 2, 3, 1).reshape(-1, Cin).contiguous().requires_grad_(True)
    y = ops.conv3x3(x_nhwc, (NB, H, W), pack, stride=stride)
    xf = x.float().requires_grad_(True)
    ref = F.conv2d(xf, conv.weight.bfloat16().float(), conv.bias, stride=stride, padding=1)
    Ho, Wo = ref.shape[2:]
    assert _rel(y, ref.permute(0, 2, 3, 1).reshape(-1, Cout)) < 6e-3
    g = torch.randn(NB, Cout, Ho, Wo, device="cuda").bfloat16()
    y.backward(g.permute(0, 2, 3, 1).reshape(-1, Cout).contiguous())
    ref.backward(g.float())
    assert _rel(x_nhwc.grad, xf.grad.permute(0, 2, 3, 1).reshape(-1, Cin)) < 8e-3


def test_linear_lora_grads():
    from flash.b200 import ops
    from flash.models.lora import LoRALinear
    torch.manual_seed(1)
    M, K, N, r = 1024, 320, 640, 64
    lin = torch.nn.Linear(K, N).cuda()
    lora = LoRALinear(lin, r, 2 * r, "gaussian").cuda()       # scaling 2
    torch.nn.init.normal_(lora.lora_B["default"].weight, std=0.05)
    for p in lin.parameters():
        p.requires_grad = False
    pack = ops.LinearPack(lora)
    x = torch.randn(M, K, device="cuda").bfloat16().requires_grad_(True)
    res = torch.randn(M, N, device="cuda").bfloat16().requires_grad_(True)
    y = ops.linear(x, pack, residual=res)
    g = torch.randn(M, N, device="cuda").bfloat16()
    y.backward(g)
    A, Bm = lora.lora_A["default"].weight, lora.lora_B["default"].weight
    xf, rf = x.detach().float().requires_grad_(True), res.detach().float().requires_grad_(True)
    Af, Bf = A.detach().clone().requires_grad_(True), Bm.detach().clone().requires_grad_(True)
    ref = xf @ lin.weight.t() + lin.bias + (xf @ Af.t()) @ Bf.t() * 2.0 + rf
    assert _rel(y, ref) < 6e-3
    ref.backward(g.float())
    assert _rel(x.grad, xf.grad) < 1e-2 and _rel(res.grad, rf.grad) < 1e-6
    assert _cos(A.grad, Af.grad) > 0.999 and _rel(A.grad, Af.grad) < 3e-2, (_cos(A.grad, Af.grad), _rel(A.grad, Af.grad))
    assert _cos(Bm.grad, Bf.grad) > 0.999 and _rel(Bm.grad, Bf.grad) < 3e-2


def test_small_unet_backward_vs_oracle():
    """LoRA gradients + input gradient of the whole B200 UNet against fp32 oracle autograd (cos >= 0.999)."""
    from test_unet_gpu import SMALL, _inputs, _pair
    prod, ora = _pair(SMALL, lora=True)
    x, t, cond = _inputs(2, 32, 32, 96, 48)
    g = torch.randn(2, 4, 32, 32, device="cuda")
    xp = x.clone().requires_grad_(True)
    out = prod(xp, t, cond)
    (out * g).sum().backward()
    xo = x.clone().requires_grad_(True)
    ref = ora(xo, t, cond)
    (ref * g).sum().backward()
    assert _rel(out, ref) < 2e-2
    assert _cos(xp.grad, xo.grad) > 0.999, _cos(xp.grad, xo.grad)
    po, pp = dict(ora.named_parameters()), dict(prod.named_parameters())
    worst = 1.0
    for n, p in pp.items():
        if "lora_" in n:
            assert p.grad is not None, n
            worst = min(worst, _cos(p.grad, po[n].grad))
        else:
            assert p.grad is None
    assert worst > 0.995, worst
    # mid-block features path (GAN backbone) back to the input
    xp2 = x.clone().requires_grad_(True)
    mid = prod(xp2, t, cond, return_intermediate=True)
    gm = torch.randn_like(mid)
    (mid * gm).sum().backward()
    xo2 = x.clone().requires_grad_(True)
    (ora(xo2, t, cond, return_intermediate=True) * gm).sum().backward()
    assert _cos(xp2.grad, xo2.grad) > 0.999


def test_sd15_like_unet_backward_vs_oracle():
    """Config-1 backbone shape (8 heads at every level -> head dims that are not 64, zero-padded packs) with the
    reference's LoRA targets: LoRA gradients and input gradient against fp32 oracle autograd."""
    from oracle.unet import SD15_KWARGS
    from test_unet_gpu import _inputs, _pair
    small15 = dict(SD15_KWARGS, block_out_channels=[320, 640, 1280, 1280], layers_per_block=1, cross_attention_dim=96)
    prod, ora = _pair(small15, lora=True, seed=3)       # head dims 40 (-> 48), 80, 160
    x, t, cond = _inputs(2, 32, 32, 96, 0)
    g = torch.randn(2, 4, 32, 32, device="cuda")
    xp, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    out, ref = prod(xp, t, cond), ora(xo, t, cond)
    assert _rel(out, ref) < 2e-2
    (out * g).sum().backward()
    (ref * g).sum().backward()
    assert _cos(xp.grad, xo.grad) > 0.999, _cos(xp.grad, xo.grad)
    po = dict(ora.named_parameters())
    worst = min(_cos(p.grad, po[n].grad) for n, p in prod.named_parameters() if "lora_" in n)
    assert worst > 0.995, worst


def test_sd15_like_distillation_training_step():
    """BASELINE config 1 pipeline at reduced width: FlashDiffusion + TrainingPipeline around an SD1.5-shaped UNet (8 heads
    per level -> padded head dims), the reference's SD1.5 discriminator on the 8x8 mid-block features."""
    from flash import recipes
    from oracle.unet import SD15_KWARGS
    small15 = dict(SD15_KWARGS, block_out_channels=[64, 128, 256, 256], cross_attention_dim=96)
    model, pipe = recipes.build_distillation(small15, recipes.sd15_discriminator(256), "cuda", lora_rank=16, K=4,
                                             conditioner=recipes.text_only_conditioner(), ucg_keys=("text_emb",), lr=1e-3)
    snap = {n: p.detach().clone() for n, p in model.named_parameters()}
    for i in range(2):
        batch = recipes.synthetic_batch(2, 64, 77, 96, 0, seed=20 + i, device="cuda", image_px=512.0)
        out = pipe.training_step(batch, i, draws={"start_idx": [1, 3][i]})
        assert float(out["loss_optimizer_0"]) > 0 and float(out["loss_optimizer_1"]) > 0
    changed = {n for n, p in model.named_parameters() if not torch.equal(p, snap[n])}
    assert any("lora_" in n for n in changed) and any(n.startswith("discriminator") for n in changed)
    assert all(("lora_" in n and n.startswith("student_denoiser")) or n.startswith("discriminator") for n in changed)
    assert all(torch.isfinite(p).all() for p in model.parameters())
