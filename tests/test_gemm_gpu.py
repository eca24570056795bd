"""GPU parity of fd_gemm (tcgen05 GEMM / implicit-GEMM conv) against plain torch fp32 math."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def raw():
    from flash.b200 import raw as r
    return r


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (4096, 640, 640), (308, 1280, 2048),
                                   (1024, 320, 1280), (16384, 1920, 640), (4, 1280, 320), (200, 72, 200)])
@pytest.mark.parametrize("bn", [0, 64, 128, 256, 512 + 128, 512 + 160, 512 + 256])
def test_gemm_plain(raw, M, N, K, bn):
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    ref = a.float() @ b.float().t()
    out = raw.gemm(a, b, force_bn=bn)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 6e-3, _rel(out, ref)
    out32 = raw.gemm(a, b, out_fp32=True, force_bn=bn)
    assert _rel(out32, ref) < 1e-5, _rel(out32, ref)


@pytest.mark.parametrize("bn", [0, 512 + 128, 512 + 160, 512 + 256])
def test_gemm_epilogue_and_lora(raw, bn):
    torch.manual_seed(0)
    M, N, K, r = 1024, 640, 640, 64
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    t = torch.randn(M, r, device="cuda").bfloat16()
    lb = (torch.randn(N, r, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    rowvec = torch.randn(4, N, device="cuda")
    ref = x.float() @ w.float().t() + t.float() @ lb.float().t() + bias + res.float() \
        + rowvec.repeat_interleave(M // 4, dim=0)
    out = raw.gemm(x, w, a2=t, b2=lb, bias=bias, residual=res, rowvec=rowvec, rows_per_group=M // 4,
                   out_fp32=True, force_bn=bn)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)


def test_gemm_geglu(raw):
    torch.manual_seed(1)
    M, C = 512, 320
    x = torch.randn(M, C, device="cuda").bfloat16()
    w = (torch.randn(8 * C, C, device="cuda") / C ** 0.5).bfloat16()   # diffusers GEGLU.proj: [value | gate]
    b = torch.randn(8 * C, device="cuda")
    res = torch.randn(M, 4 * C, device="cuda").bfloat16()
    h = x.float() @ w.float().t() + b
    val, gate = h.chunk(2, dim=-1)
    ref = val * torch.nn.functional.gelu(gate) + res.float()
    # interleave 16 value rows / 16 gate rows
    wv, wg = w[:4 * C].view(-1, 16, C), w[4 * C:].view(-1, 16, C)
    wi = torch.stack([wv, wg], dim=1).reshape(8 * C, C).contiguous()
    bi = torch.stack([b[:4 * C].view(-1, 16), b[4 * C:].view(-1, 16)], dim=1).reshape(-1).contiguous()
    out = raw.gemm(x, wi, bias=bi, geglu=True, residual=res, out_fp32=True)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(2, 32, 32, 64, 64), (1, 64, 64, 320, 320), (2, 128, 128, 64, 128),
                                             (4, 8, 8, 128, 192), (1, 16, 16, 1920, 640), (2, 32, 32, 8, 320)])
@pytest.mark.parametrize("bn", [0, 128, 512 + 128, 512 + 160, 512 + 256])
def test_conv3x3(raw, NB, H, W, Cin, Cout, bn):
    torch.manual_seed(NB + H + Cin)
    x = torch.randn(NB, Cin, H, W, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda")
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.functional.conv2d(x.float(), w.float(), bias, padding=1)     # NCHW
    ref = ref.permute(0, 2, 3, 1).reshape(NB * H * W, Cout)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    cpad = (Cin + 63) // 64 * 64
    wp = torch.zeros(Cout, 3, 3, cpad, device="cuda", dtype=torch.bfloat16)
    wp[..., :Cin] = w.permute(0, 2, 3, 1)
    wp = wp.reshape(Cout, 9 * cpad).contiguous()
    out = raw.gemm(x_nhwc, wp, bias=bias, out_fp32=True, M=NB * H * W,
                   conv=dict(NB_in=NB, H=H, W=W, C=Cin, taps=raw.TAPS_3X3), force_bn=bn)
    assert _rel(out, ref) < 3e-5, _rel(out, ref)


def test_gemm_ragged_k(raw):
    torch.manual_seed(5)
    M, N, K = 640, 192, 154                      # K not a multiple of 8: row strides padded to 160
    abuf = torch.randn(M, 160, device="cuda").bfloat16()
    bbuf = torch.randn(N, 160, device="cuda").bfloat16()
    a, b = abuf[:, :K], bbuf[:, :K]
    out = raw.gemm(a, b, out_fp32=True)
    assert _rel(out, a.float() @ b.float().t()) < 1e-5


@pytest.mark.parametrize("bn", [0, 128, 512 + 160, 512 + 256])
@pytest.mark.parametrize("geglu", [False, True])
def test_gemm_layernorm_fold_and_rowstats(raw, bn, geglu):
    """LayerNorm(x) W^T + b via the folded epilogue == explicit LayerNorm then GEMM; producer row statistics."""
    torch.manual_seed(7)
    M, C, N = 1024, 640, 1280
    h0 = torch.randn(M, C, device="cuda").bfloat16()
    w0 = (torch.randn(C, C, device="cuda") / C ** 0.5).bfloat16()
    res = (torch.randn(M, C, device="cuda") * 2 + 0.7).bfloat16()
    stats = torch.empty(M, 2, device="cuda")
    x = raw.gemm(h0, w0, residual=res, rowstats=stats, force_bn=bn)          # producer: residual stream + its stats
    xf = x.float()
    assert _rel(stats[:, 0], xf.sum(1)) < 1e-4 and _rel(stats[:, 1], (xf * xf).sum(1)) < 1e-4
    gamma, beta = torch.randn(C, device="cuda") * 0.5 + 1.0, torch.randn(C, device="cuda") * 0.3
    w = torch.randn(N, C, device="cuda") / C ** 0.5
    b = torch.randn(N, device="cuda")
    ref = torch.nn.functional.layer_norm(xf, (C,), gamma, beta, 1e-5) @ w.t() + b
    wg = (w * gamma[None, :])
    bias = b + w @ beta
    if geglu:
        val, gate = ref.chunk(2, dim=-1)
        ref = val * torch.nn.functional.gelu(gate)
        wv, wgt = wg[:N // 2].view(-1, 16, C), wg[N // 2:].view(-1, 16, C)
        wg = torch.stack([wv, wgt], dim=1).reshape(N, C)
        bias = torch.stack([bias[:N // 2].view(-1, 16), bias[N // 2:].view(-1, 16)], dim=1).reshape(-1)
    wp = wg.contiguous().bfloat16()
    colsum = wp.float().sum(1).contiguous()
    out = raw.gemm(x, wp, bias=bias.contiguous(), geglu=geglu, ln=(stats, colsum, C, 1e-5), out_fp32=True, force_bn=bn)
    assert _rel(out, ref) < 6e-3, _rel(out, ref)


# ------------------------------------------------------------------------------------------ work split + TMA store
SPLIT_MODES = {"auto": 0, "streamk": 1024, "whole_tiles": 2048}


@pytest.mark.parametrize("M,N,K", [(8192, 1280, 1280),    # 160 tiles on 74 pairs: the in-situ shape (partial wave of 12 tiles)
                                   (20000, 640, 1024),    # 79 x 3 tiles: ragged M, stream-K region of 1.2 waves
                                   (4096, 640, 640),      # 48 tiles < 74 pairs: every tile cut into column slices
                                   (616, 2560, 2048),     # 77-token K/V projection of the 2B teacher batch: ragged M
                                   (1000, 200, 1000)])    # ragged everything (TMA store clips rows and columns)
@pytest.mark.parametrize("bn", [0, 512 + 128, 512 + 160, 512 + 256])
@pytest.mark.parametrize("mode", ["auto", "streamk", "whole_tiles"])
def test_gemm_work_split_epilogues(raw, M, N, K, bn, mode):
    """column slices of the partial wave / hybrid stream-K partial-tile exchange / whole tiles, each with the TMA-store
    epilogue and every fused epilogue term, bf16 and fp32 outputs, repeated launches (the stream-K flags must be left
    clean)."""
    if bn == 0 and mode != "auto":
        pytest.skip("work-split override needs an explicit tile")
    fb = bn | SPLIT_MODES[mode]
    torch.manual_seed(M * 7 + N)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    ref = a.float() @ b.float().t() + bias + res.float()
    for _ in range(3):
        out = raw.gemm(a, b, bias=bias, residual=res, force_bn=fb)
        assert _rel(out, ref) < 6e-3, _rel(out, ref)
    out32 = raw.gemm(a, b, bias=bias, residual=res, out_fp32=True, force_bn=fb)
    assert _rel(out32, ref) < 1e-5, _rel(out32, ref)
    # row statistics of the stored bf16 output (LayerNorm-fold producer side)
    stats = torch.empty(M, 2, device="cuda")
    out = raw.gemm(a, b, bias=bias, residual=res, rowstats=stats, force_bn=fb)
    of = out.float()
    assert torch.allclose(stats[:, 0], of.sum(1), rtol=1e-3, atol=1e-2)
    assert torch.allclose(stats[:, 1], (of * of).sum(1), rtol=1e-3, atol=1e-2)
    # strided output view (fused qkv buffers are written through column slices)
    big = torch.zeros(M, N + 64, device="cuda", dtype=torch.bfloat16)
    raw.gemm(a, b, bias=bias, residual=res, out=big[:, 32:32 + N], force_bn=fb)
    assert _rel(big[:, 32:32 + N], ref) < 6e-3
    assert float(big[:, :32].abs().max()) == 0 and float(big[:, 32 + N:].abs().max()) == 0
    ws = raw.gemm_workspace(a.device)
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0          # flags reset by the readers


@pytest.mark.parametrize("mode", ["auto", "streamk", "whole_tiles"])
def test_gemm_work_split_geglu_lora(raw, mode):
    torch.manual_seed(3)
    M, C, r = 20480, 640, 64
    x = torch.randn(M, C, device="cuda").bfloat16()
    w = (torch.randn(2 * C, C, device="cuda") / C ** 0.5).bfloat16()
    b = torch.randn(2 * C, device="cuda")
    h = x.float() @ w.float().t() + b
    val, gate = h.chunk(2, dim=-1)
    ref = val * torch.nn.functional.gelu(gate)
    wv, wg = w[:C].view(-1, 16, C), w[C:].view(-1, 16, C)
    wi = torch.stack([wv, wg], dim=1).reshape(2 * C, C).contiguous()
    bi = torch.stack([b[:C].view(-1, 16), b[C:].view(-1, 16)], dim=1).reshape(-1).contiguous()
    for bn in (512 + 128, 512 + 256):
        out = raw.gemm(x, wi, bias=bi, geglu=True, force_bn=bn | SPLIT_MODES[mode])     # 32-byte-row TMA store
        assert _rel(out, ref) < 6e-3, (bn, _rel(out, ref))
    # second K segment (LoRA) across a stream-K cut / inside column slices
    t = torch.randn(M, r, device="cuda").bfloat16()
    w2 = (torch.randn(1280, C, device="cuda") / C ** 0.5).bfloat16()
    lb = (torch.randn(1280, r, device="cuda") * 0.05).bfloat16()
    ref2 = x.float() @ w2.float().t() + t.float() @ lb.float().t()
    for bn in (512 + 128, 512 + 160, 512 + 256):
        out = raw.gemm(x, w2, a2=t, b2=lb, force_bn=bn | SPLIT_MODES[mode])
        assert _rel(out, ref2) < 6e-3, (bn, _rel(out, ref2))


@pytest.mark.parametrize("mode", ["auto", "streamk"])
def test_conv3x3_work_split(raw, mode):
    """implicit-GEMM conv through the same work splits (4-D TMA A operand, K = 9 taps x channels)."""
    torch.manual_seed(5)
    NB, H, W, Cin, Cout = 5, 64, 64, 128, 320             # M = 20480 rows: 80 x 2 tiles at BN = 160/256
    x = torch.randn(NB, Cin, H, W, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda")
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.functional.conv2d(x.float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(NB * H * W, Cout)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    cpad = (Cin + 63) // 64 * 64
    wp = torch.zeros(Cout, 9, cpad, device="cuda")
    wp[:, :, :Cin] = w.float().permute(0, 2, 3, 1).reshape(Cout, 9, Cin)
    wp = wp.reshape(Cout, 9 * cpad).bfloat16()
    for bn in (512 + 128, 512 + 160, 512 + 256):
        out = raw.gemm(x_nhwc, wp, bias=bias, conv=dict(NB_in=NB, H=H, W=W, C=Cin, taps=raw.TAPS_3X3), M=NB * H * W,
                       force_bn=bn | SPLIT_MODES[mode])
        assert _rel(out, ref) < 6e-3, (bn, _rel(out, ref))


@pytest.mark.parametrize("mode", ["auto", "streamk", "whole_tiles"])
@pytest.mark.parametrize("NB,HW,N,K,bn", [(4, 1024, 1280, 1280, 0), (3, 4096, 320, 640, 512 + 160),
                                          (8, 64, 1280, 320, 0), (2, 2048, 96, 200, 512 + 128),
                                          (5, 4096, 640, 1920, 512 + 256)])
def test_gemm_colstats(raw, NB, HW, N, K, bn, mode):
    """FdGemmArgs.colstats_out: per-image column (sum, sum of squares) of the STORED bf16 output, accumulated by the
    epilogue (GroupNorm statistics from the producer) — with bias + residual, through every work split."""
    torch.manual_seed(NB + HW + N)
    M = NB * HW
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    ref = raw.gemm(a, b, bias=bias, residual=res, force_bn=bn | SPLIT_MODES[mode])
    cs = torch.zeros(NB, N, 2, device="cuda")
    out = raw.gemm(a, b, bias=bias, residual=res, force_bn=bn | SPLIT_MODES[mode], colstats=cs)
    assert torch.equal(out, ref)                         # the statistics do not perturb the output
    o = out.float().view(NB, HW, N)
    want = torch.stack([o.sum(1), (o * o).sum(1)], dim=-1)
    scale = (o * o).sum(1).max().item()
    assert (cs - want).abs().max().item() < 2e-5 * scale + 1e-2, (cs - want).abs().max().item()
    # accumulate semantics: a second call adds on top
    raw.gemm(a, b, bias=bias, residual=res, force_bn=bn | SPLIT_MODES[mode], colstats=cs)
    assert (cs - 2 * want).abs().max().item() < 4e-5 * scale + 2e-2


def test_conv3x3_colstats(raw):
    torch.manual_seed(9)
    NB, H, W, Cin, Cout = 3, 32, 32, 128, 320
    x = torch.randn(NB, H, W, Cin, device="cuda").bfloat16()
    cpad = (Cin + 63) // 64 * 64
    wp = (torch.randn(Cout, 9 * cpad, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda")
    conv = dict(NB_in=NB, H=H, W=W, C=Cin, taps=raw.TAPS_3X3)
    ref = raw.gemm(x, wp, bias=bias, conv=conv, M=NB * H * W)
    cs = torch.zeros(NB, Cout, 2, device="cuda")
    out = raw.gemm(x, wp, bias=bias, conv=conv, M=NB * H * W, colstats=cs)
    assert torch.equal(out, ref)
    o = out.float().view(NB, H * W, Cout)
    want = torch.stack([o.sum(1), (o * o).sum(1)], dim=-1)
    assert (cs - want).abs().max().item() < 2e-5 * (o * o).sum(1).max().item() + 1e-2


def test_gemm_colstats_rejects_unsupported(raw):
    a = torch.randn(256, 64, device="cuda").bfloat16()
    b = torch.randn(72, 64, device="cuda").bfloat16()          # N % 32 != 0
    with pytest.raises(RuntimeError):
        raw.gemm(a, b, colstats=torch.zeros(2, 72, 2, device="cuda"))
