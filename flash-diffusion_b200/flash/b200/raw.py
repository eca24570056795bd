"""Raw (non-autograd) Python entry points over the C ABI: tensors in, tensors out.

Every function launches hand-written sm_100a kernels from libflashb200.so on the current CUDA
stream.  Tensors must be CUDA tensors; bf16 activations are channels-last ([rows, C] / NHWC).
"""
from ctypes import byref, c_float, c_int32, c_int64

import torch

from . import lib as _l
from .lib import FdGemmArgs, check, load, ptr, stream_ptr

BF16 = torch.bfloat16

TAPS_3X3 = [(0, kh - 1, kw - 1) for kh in range(3) for kw in range(3)]


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name}: B200 backend needs CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")


_GEMM_WS = {}


def gemm_workspace(device):
    """Caller-owned stream-K scratch of fd_gemm (include/flashb200.h): one zero-filled buffer per device.  Every GEMM
    of a device runs stream-ordered on the compute stream (eager launches and CUDA-graph replays alike), so one
    buffer is enough; it is created on the first call, which must not happen under stream capture."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _GEMM_WS.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("fd_gemm workspace must be created before CUDA-graph capture (run one eager GEMM first)")
        ws = torch.zeros(int(load().fd_gemm_workspace_bytes()), dtype=torch.uint8, device=device)
        _GEMM_WS[key] = ws
    return ws


def gemm(a1, b1, *, a2=None, b2=None, bias=None, rowvec=None, rows_per_group=0, geglu=False,
         residual=None, out=None, out_fp32=False, conv=None, M=None, force_bn=0, ln=None, rowstats=None,
         act=0, rowscale=None, rows_per_group_scale=0, rowstats_prezeroed=False, colstats=None):
    """acc = a1 @ b1.T (+ a2 @ b2.T) with the fused epilogue of fd_gemm (include/flashb200.h).

    colstats: ZERO-FILLED [images, N, 2] fp32 — per-image column (sum, sum of squares) of the stored output, the
    statistics `groupnorm_apply_cols` of the following GroupNorm consumes (FdGemmArgs.colstats_out).

    a1: [M, K1] bf16 (or, with conv=dict(NB_in,H,W,C,taps), an NHWC tensor), b1: [N, K1] bf16.
    """
    lib = load()
    _req(a1, BF16, "a1"); _req(b1, BF16, "b1"); _req(a2, BF16, "a2"); _req(b2, BF16, "b2")
    _req(bias, torch.float32, "bias"); _req(rowvec, torch.float32, "rowvec")
    _req(residual, BF16, "residual")
    args = FdGemmArgs()
    N, K1 = b1.shape
    assert b1.stride(1) == 1
    if conv is None:
        assert a1.dim() == 2 and a1.stride(1) == 1 and a1.shape[1] == K1, (a1.shape, b1.shape)
        M = a1.shape[0]
        args.lda1 = a1.stride(0)
    else:
        assert a1.is_contiguous()
        assert M is not None
        taps = conv["taps"]
        args.conv_taps = len(taps)
        args.NB_in, args.H, args.W, args.C = conv["NB_in"], conv["H"], conv["W"], conv["C"]
        for i, (dn, dh, dw) in enumerate(taps):
            args.tap_dn[i], args.tap_dh[i], args.tap_dw[i] = dn, dh, dw
    args.M, args.N, args.K1 = M, N, K1
    args.a1, args.b1, args.ldb1 = ptr(a1), ptr(b1), b1.stride(0)
    if a2 is not None:
        assert a2.shape[0] == M and a2.stride(1) == 1 and b2.stride(1) == 1
        assert b2.shape == (N, a2.shape[1])
        args.a2, args.lda2, args.b2, args.ldb2, args.K2 = ptr(a2), a2.stride(0), ptr(b2), b2.stride(0), a2.shape[1]
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a1.device, dtype=torch.float32 if out_fp32 else BF16)
    else:
        assert out.shape == (M, n_out) and out.stride(1) == 1
        assert out.dtype == (torch.float32 if out_fp32 else BF16)
    if bias is not None:
        assert bias.numel() == N and bias.is_contiguous()
        args.bias = ptr(bias)
    if rowvec is not None:
        assert rowvec.dim() == 2 and rowvec.stride(1) == 1 and rowvec.shape[1] == N
        args.rowvec, args.rows_per_group, args.ldrv = ptr(rowvec), rows_per_group, rowvec.stride(0)
    args.geglu = 1 if geglu else 0
    if residual is not None:
        assert residual.shape == (M, n_out) and residual.stride(1) == 1
        args.residual, args.ldr = ptr(residual), residual.stride(0)
    args.out, args.ldo, args.out_fp32 = ptr(out), out.stride(0), 1 if out_fp32 else 0
    args.force_bn = force_bn
    if ln is not None:
        # LayerNorm fold: ln = (row_stats [M,2] fp32 raw sums, colsum [N] fp32, channels C, eps)
        st, colsum, C, eps = ln
        assert st.dtype == torch.float32 and st.shape == (M, 2) and st.is_contiguous()
        assert colsum.dtype == torch.float32 and colsum.numel() == N and colsum.is_contiguous()
        args.ln_stats, args.ln_colsum, args.ln_inv_c, args.ln_eps = ptr(st), ptr(colsum), 1.0 / C, eps
    args.act = act
    if rowscale is not None:
        # AdaLN gate: acc *= rowscale[row // rows_per_group_scale, :]
        assert rowscale.dtype == torch.float32 and rowscale.dim() == 2 and rowscale.stride(1) == 1
        assert rowscale.shape[1] == n_out and rows_per_group_scale > 0
        args.rowscale, args.rows_per_group_scale, args.ldrs = ptr(rowscale), rows_per_group_scale, rowscale.stride(0)
    if rowstats is not None:
        assert rowstats.dtype == torch.float32 and rowstats.shape == (M, 2) and rowstats.is_contiguous()
        args.rowstats_out = ptr(rowstats)
        args.rowstats_prezeroed = 1 if rowstats_prezeroed else 0
    if colstats is not None:
        assert colstats.dtype == torch.float32 and colstats.is_contiguous() and colstats.dim() == 3
        assert colstats.shape[1:] == (N, 2) and M % colstats.shape[0] == 0, (colstats.shape, M, N)
        args.colstats_out, args.colstats_rows = ptr(colstats), M // colstats.shape[0]
    ws = gemm_workspace(a1.device)
    args.workspace, args.workspace_bytes = ptr(ws), ws.numel()
    check(lib.fd_gemm(byref(args), stream_ptr()), "fd_gemm")
    return out


# ------------------------------------------------------------------ normalisation
def groupnorm_stats(x, NB, HW, C, G, eps):
    lib = load(); _req(x, BF16, "x")
    stats = torch.empty((NB, G, 2), device=x.device, dtype=torch.float32)
    check(lib.fd_groupnorm_stats(ptr(x), ptr(stats), c_int32(NB), c_int32(HW), c_int32(C), c_int32(G),
                                 c_float(eps), stream_ptr()), "fd_groupnorm_stats")
    return stats


def groupnorm_apply(x, stats, gamma, beta, NB, HW, C, G, silu):
    lib = load(); _req(x, BF16, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    y = torch.empty_like(x)
    check(lib.fd_groupnorm_apply(ptr(x), ptr(stats), ptr(gamma), ptr(beta), ptr(y), c_int32(NB), c_int32(HW),
                                 c_int32(C), c_int32(G), c_int32(1 if silu else 0), stream_ptr()),
          "fd_groupnorm_apply")
    return y


def groupnorm_fwd(x, gamma, beta, NB, HW, C, G, eps, silu, want_stats=False):
    """GroupNorm(+SiLU) forward in two launches (fd_groupnorm_fwd); returns y or (y, stats[NB,G,2] = mean, rstd)."""
    lib = load(); _req(x, BF16, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    y = torch.empty_like(x)
    raw_sums = torch.empty((NB, G, 2), device=x.device, dtype=torch.float32)
    stats = torch.empty((NB, G, 2), device=x.device, dtype=torch.float32) if want_stats else None
    check(lib.fd_groupnorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(raw_sums), ptr(stats), c_int32(NB),
                               c_int32(HW), c_int32(C), c_int32(G), c_float(eps), c_int32(1 if silu else 0),
                               stream_ptr()), "fd_groupnorm_fwd")
    return (y, stats) if want_stats else y


def groupnorm_apply_cols(x, colstats, gamma, beta, NB, HW, C, G, eps, silu):
    """GroupNorm(+SiLU) forward in ONE launch from the per-image column sums [NB, C, 2] the producing GEMM / conv left
    (gemm(colstats=...)): no reduction pass over x."""
    lib = load(); _req(x, BF16, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    assert colstats.dtype == torch.float32 and colstats.shape == (NB, C, 2) and colstats.is_contiguous()
    y = torch.empty_like(x)
    check(lib.fd_groupnorm_apply_cols(ptr(x), ptr(colstats), ptr(gamma), ptr(beta), ptr(y), None, c_int32(NB),
                                      c_int32(HW), c_int32(C), c_int32(G), c_float(eps), c_int32(1 if silu else 0),
                                      stream_ptr()), "fd_groupnorm_apply_cols")
    return y


def groupnorm_bwd(x, stats, gamma, beta, dy, NB, HW, C, G, silu):
    lib = load(); _req(x, BF16, "x"); _req(dy, BF16, "dy")
    dx = torch.empty_like(x)
    scratch = torch.empty((NB, G, 2), device=x.device, dtype=torch.float32)
    check(lib.fd_groupnorm_bwd(ptr(x), ptr(stats), ptr(gamma), ptr(beta), ptr(dy), ptr(dx), ptr(scratch),
                               c_int32(NB), c_int32(HW), c_int32(C), c_int32(G), c_int32(1 if silu else 0),
                               stream_ptr()), "fd_groupnorm_bwd")
    return dx


def layernorm_fwd(x, gamma, beta, eps, save_stats=False):
    lib = load(); _req(x, BF16, "x")
    rows, C = x.shape
    assert x.is_contiguous()
    y = torch.empty_like(x)
    stats = torch.empty((rows, 2), device=x.device, dtype=torch.float32) if save_stats else None
    check(lib.fd_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(stats), c_int32(rows), c_int32(C),
                               c_float(eps), stream_ptr()), "fd_layernorm_fwd")
    return (y, stats) if save_stats else y


def layernorm_modulate(x, scale, shift, rows_per_batch, eps):
    """y = LayerNorm(x) * (1 + scale[b]) + shift[b]; scale/shift fp32 [B, C] views with equal row stride."""
    lib = load(); _req(x, BF16, "x"); _req(scale, torch.float32, "scale"); _req(shift, torch.float32, "shift")
    rows, C = x.shape
    assert x.is_contiguous() and scale.stride(1) == 1 and shift.stride(1) == 1 and scale.stride(0) == shift.stride(0)
    y = torch.empty_like(x)
    check(lib.fd_layernorm_modulate(ptr(x), ptr(scale), ptr(shift), c_int64(scale.stride(0)), ptr(y), c_int32(rows),
                                    c_int32(C), c_int32(rows_per_batch), c_float(eps), stream_ptr()),
          "fd_layernorm_modulate")
    return y


def layernorm_modulate_bwd(x, dy, scale, rows_per_batch, eps):
    """-> dx [rows, C] bf16, dscale [B, C] fp32, dshift [B, C] fp32 (see fd_layernorm_modulate_bwd)."""
    lib = load(); _req(x, BF16, "x"); _req(dy, BF16, "dy"); _req(scale, torch.float32, "scale")
    rows, C = x.shape
    assert x.is_contiguous() and dy.is_contiguous() and scale.stride(1) == 1 and rows % rows_per_batch == 0
    B = rows // rows_per_batch
    dx = torch.empty_like(x)
    dscale = torch.empty((B, C), device=x.device, dtype=torch.float32)
    dshift = torch.empty((B, C), device=x.device, dtype=torch.float32)
    check(lib.fd_layernorm_modulate_bwd(ptr(x), ptr(dy), ptr(scale), c_int64(scale.stride(0)), ptr(dx), ptr(dscale),
                                        ptr(dshift), c_int32(rows), c_int32(C), c_int32(rows_per_batch), c_float(eps),
                                        stream_ptr()), "fd_layernorm_modulate_bwd")
    return dx, dscale, dshift


def gate_residual(h, gate, res, rows_per_batch):
    """out = res + gate[b] * h; gate fp32 [B, C] view (unit column stride)."""
    lib = load(); _req(h, BF16, "h"); _req(res, BF16, "res"); _req(gate, torch.float32, "gate")
    rows, C = h.shape
    assert h.is_contiguous() and res.is_contiguous() and res.shape == h.shape and gate.stride(1) == 1
    out = torch.empty_like(h)
    check(lib.fd_gate_residual(ptr(h), ptr(gate), c_int64(gate.stride(0)), ptr(res), ptr(out), c_int32(rows),
                               c_int32(C), c_int32(rows_per_batch), stream_ptr()), "fd_gate_residual")
    return out


def gate_bwd(dout, h, gate, rows_per_batch):
    """-> dh [rows, C] bf16 = gate[b] * dout, dgate [B, C] fp32 = per-sample column sums of dout * h."""
    lib = load(); _req(dout, BF16, "dout"); _req(h, BF16, "h"); _req(gate, torch.float32, "gate")
    rows, C = h.shape
    assert h.is_contiguous() and dout.is_contiguous() and gate.stride(1) == 1 and rows % rows_per_batch == 0
    dh = torch.empty_like(h)
    dgate = torch.empty((rows // rows_per_batch, C), device=h.device, dtype=torch.float32)
    check(lib.fd_gate_bwd(ptr(dout), ptr(h), ptr(gate), c_int64(gate.stride(0)), ptr(dh), ptr(dgate), c_int32(rows),
                          c_int32(C), c_int32(rows_per_batch), stream_ptr()), "fd_gate_bwd")
    return dh, dgate


def gelu_tanh_bwd(acc, dout):
    lib = load(); _req(acc, BF16, "acc"); _req(dout, BF16, "dout")
    assert acc.is_contiguous() and dout.is_contiguous() and acc.shape == dout.shape and acc.numel() % 8 == 0
    dacc = torch.empty_like(acc)
    check(lib.fd_gelu_tanh_bwd(ptr(acc), ptr(dout), ptr(dacc), c_int64(acc.numel()), stream_ptr()), "fd_gelu_tanh_bwd")
    return dacc


def patchify(dy, h, w, p, Cout):
    """gradient of unpatchify: dy [NB, Ckeep, h*p, w*p] fp32 -> [NB*h*w, p*p*Cout] bf16"""
    lib = load(); _req(dy, torch.float32, "dy")
    NB, Ckeep = dy.shape[:2]
    assert dy.is_contiguous() and dy.shape[2:] == (h * p, w * p)
    dx = torch.empty((NB * h * w, p * p * Cout), device=dy.device, dtype=BF16)
    check(lib.fd_patchify(ptr(dy), ptr(dx), c_int32(NB), c_int32(h), c_int32(w), c_int32(p), c_int32(Cout),
                          c_int32(Ckeep), stream_ptr()), "fd_patchify")
    return dx


def unpatchify(x, NB, h, w, p, Cout, Ckeep):
    """x [NB*h*w, p*p*Cout] fp32 -> [NB, Ckeep, h*p, w*p] fp32"""
    lib = load(); _req(x, torch.float32, "x")
    assert x.is_contiguous() and x.shape == (NB * h * w, p * p * Cout)
    y = torch.empty((NB, Ckeep, h * p, w * p), device=x.device, dtype=torch.float32)
    check(lib.fd_unpatchify(ptr(x), ptr(y), c_int32(NB), c_int32(h), c_int32(w), c_int32(p), c_int32(Cout),
                            c_int32(Ckeep), stream_ptr()), "fd_unpatchify")
    return y


def layernorm_bwd(x, stats, gamma, dy):
    lib = load(); _req(x, BF16, "x"); _req(dy, BF16, "dy")
    rows, C = x.shape
    assert x.is_contiguous() and dy.is_contiguous()
    dx = torch.empty_like(x)
    check(lib.fd_layernorm_bwd(ptr(x), ptr(stats), ptr(gamma), ptr(dy), ptr(dx), c_int32(rows), c_int32(C),
                               stream_ptr()), "fd_layernorm_bwd")
    return dx


# ------------------------------------------------------------------ attention
def attention_fwd(q, k, v, H, scale=None, need_lse=False, head_dim=64, kv_len=None):
    """q [B,Nq,H*d] view (last-dim stride 1), k/v [B,Nkv,H*d] views -> o [B,Nq,H*d] bf16.
    d = 64 runs the tuned kernel (fd_attn_fwd); any other multiple of 16 up to 192, or a key-padding mask
    kv_len [B] int32, runs fd_attn_fwd_generic."""
    lib = load(); _req(q, BF16, "q"); _req(k, BF16, "k"); _req(v, BF16, "v")
    B, Nq, HD = q.shape
    Nkv = k.shape[1]
    assert HD == H * head_dim and q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    o = torch.empty((B, Nq, HD), device=q.device, dtype=BF16)
    lse = torch.empty((B, H, Nq), device=q.device, dtype=torch.float32) if need_lse else None
    a = _l.FdAttnArgs()
    a.q, a.ldq, a.q_batch_stride = ptr(q), q.stride(1), q.stride(0)
    a.k, a.ldk, a.k_batch_stride = ptr(k), k.stride(1), k.stride(0)
    a.v, a.ldv, a.v_batch_stride = ptr(v), v.stride(1), v.stride(0)
    a.o, a.ldo, a.o_batch_stride = ptr(o), o.stride(1), o.stride(0)
    a.lse = ptr(lse)
    a.B, a.H, a.Nq, a.Nkv = B, H, Nq, Nkv
    a.scale = scale if scale is not None else head_dim ** -0.5
    if head_dim == 64 and kv_len is None:
        check(lib.fd_attn_fwd(byref(a), stream_ptr()), "fd_attn_fwd")
    else:
        if kv_len is not None:
            assert kv_len.dtype == torch.int32 and kv_len.numel() == B and kv_len.is_cuda
        check(lib.fd_attn_fwd_generic(byref(a), c_int32(head_dim), ptr(kv_len), stream_ptr()), "fd_attn_fwd_generic")
    return (o, lse) if need_lse else o


def attention_bwd(q, k, v, o, lse, do, H, scale=None, dq=None, dk=None, dv=None, head_dim=64, kv_len=None):
    """Gradients of attention_fwd.  dq/dk/dv may be given as (strided) output views, e.g. slices of one fused
    [B, N, 3*H*d] buffer; otherwise contiguous tensors are allocated.  d = 64 without a mask runs the tuned kernel
    (fd_attn_bwd), everything else fd_attn_bwd_generic."""
    lib = load(); _req(do, BF16, "do")
    B, Nq, HD = q.shape
    Nkv = k.shape[1]
    assert HD == H * head_dim and do.stride(2) == 1 and o.stride(2) == 1
    dq = torch.empty((B, Nq, HD), device=q.device, dtype=BF16) if dq is None else dq
    dk = torch.empty((B, Nkv, HD), device=q.device, dtype=BF16) if dk is None else dk
    dv = torch.empty((B, Nkv, HD), device=q.device, dtype=BF16) if dv is None else dv
    delta = torch.empty((B, H, Nq), device=q.device, dtype=torch.float32)
    if head_dim <= 80:
        dq_accum = torch.empty((B, Nq, HD), device=q.device, dtype=torch.float32)
    else:       # short-sequence kernels: scratch for P and dS
        dq_accum = torch.empty((2 * B * H * Nq * Nkv,), device=q.device, dtype=torch.float32)
    a = _l.FdAttnBwdArgs()
    f = a.f
    f.q, f.ldq, f.q_batch_stride = ptr(q), q.stride(1), q.stride(0)
    f.k, f.ldk, f.k_batch_stride = ptr(k), k.stride(1), k.stride(0)
    f.v, f.ldv, f.v_batch_stride = ptr(v), v.stride(1), v.stride(0)
    f.o, f.ldo, f.o_batch_stride = ptr(o), o.stride(1), o.stride(0)
    f.lse = ptr(lse)
    f.B, f.H, f.Nq, f.Nkv = B, H, Nq, Nkv
    f.scale = scale if scale is not None else head_dim ** -0.5
    a.d_o, a.lddo, a.do_batch_stride = ptr(do), do.stride(1), do.stride(0)
    a.dq, a.lddq, a.dq_batch_stride = ptr(dq), dq.stride(1), dq.stride(0)
    a.dk, a.lddk, a.dk_batch_stride = ptr(dk), dk.stride(1), dk.stride(0)
    a.dv, a.lddv, a.dv_batch_stride = ptr(dv), dv.stride(1), dv.stride(0)
    a.delta, a.dq_accum = ptr(delta), ptr(dq_accum)
    if head_dim == 64 and kv_len is None:
        check(lib.fd_attn_bwd(byref(a), stream_ptr()), "fd_attn_bwd")
    else:
        if kv_len is not None:
            assert kv_len.dtype == torch.int32 and kv_len.numel() == B and kv_len.is_cuda
        check(lib.fd_attn_bwd_generic(byref(a), c_int32(head_dim), ptr(kv_len), stream_ptr()), "fd_attn_bwd_generic")
    return dq, dk, dv


def softmax_rows(x, scale=1.0):
    """bf16 [rows, L] -> softmax(scale * x) over the last dim (fd_softmax_rows)."""
    lib = load(); _req(x, BF16, "x")
    assert x.dim() == 2 and x.stride(1) == 1
    y = torch.empty((x.shape[0], x.shape[1]), device=x.device, dtype=BF16)
    check(lib.fd_softmax_rows(ptr(x), c_int64(x.stride(0)), ptr(y), c_int64(y.stride(0)), c_int32(x.shape[0]),
                              c_int32(x.shape[1]), c_float(scale), stream_ptr()), "fd_softmax_rows")
    return y


def softmax_rows_bwd(p, dp, scale=1.0):
    lib = load(); _req(p, BF16, "p"); _req(dp, BF16, "dp")
    assert p.is_contiguous() and dp.is_contiguous() and p.shape == dp.shape
    ds = torch.empty_like(p)
    check(lib.fd_softmax_rows_bwd(ptr(p), ptr(dp), ptr(ds), c_int64(p.stride(0)), c_int32(p.shape[0]),
                                  c_int32(p.shape[1]), c_float(scale), stream_ptr()), "fd_softmax_rows_bwd")
    return ds


def maxpool2x2(x, NB, H, W, C):
    lib = load(); _req(x, BF16, "x")
    assert x.is_contiguous()
    y = torch.empty((NB * (H // 2) * (W // 2), C), device=x.device, dtype=BF16)
    check(lib.fd_maxpool2x2(ptr(x), ptr(y), c_int32(NB), c_int32(H), c_int32(W), c_int32(C), stream_ptr()), "fd_maxpool2x2")
    return y


def maxpool2x2_bwd(x, dy, NB, H, W, C):
    lib = load(); _req(x, BF16, "x"); _req(dy, BF16, "dy")
    assert x.is_contiguous() and dy.is_contiguous()
    dx = torch.empty_like(x)
    check(lib.fd_maxpool2x2_bwd(ptr(x), ptr(dy), ptr(dx), c_int32(NB), c_int32(H), c_int32(W), c_int32(C), stream_ptr()),
          "fd_maxpool2x2_bwd")
    return dx


def relu_bwd(y, dy):
    lib = load(); _req(y, BF16, "y"); _req(dy, BF16, "dy")
    assert y.is_contiguous() and dy.is_contiguous() and y.shape == dy.shape
    dx = torch.empty_like(y)
    check(lib.fd_relu_bwd(ptr(y), ptr(dy), ptr(dx), c_int64(y.numel()), stream_ptr()), "fd_relu_bwd")
    return dx


def lpips_layer(f0, f1, w, out, NB, HW, C):
    """out[n] += LPIPS distance of one feature layer (out fp32 [NB], accumulated in place)."""
    lib = load(); _req(f0, BF16, "f0"); _req(f1, BF16, "f1"); _req(w, torch.float32, "w"); _req(out, torch.float32, "out")
    assert f0.is_contiguous() and f1.is_contiguous() and f0.shape == f1.shape and w.numel() == C and out.numel() == NB
    check(lib.fd_lpips_layer(ptr(f0), ptr(f1), ptr(w), ptr(out), c_int32(NB), c_int32(HW), c_int32(C), stream_ptr()),
          "fd_lpips_layer")
    return out


def lpips_layer_bwd(f0, f1, w, gout, NB, HW, C):
    lib = load(); _req(f0, BF16, "f0"); _req(f1, BF16, "f1"); _req(gout, torch.float32, "gout")
    df0 = torch.empty_like(f0)
    check(lib.fd_lpips_layer_bwd(ptr(f0), ptr(f1), ptr(w), ptr(gout.contiguous()), ptr(df0), c_int32(NB), c_int32(HW),
                                 c_int32(C), stream_ptr()), "fd_lpips_layer_bwd")
    return df0


# ------------------------------------------------------------------ layout / elementwise
def nchw_to_nhwc(x, Cpad):
    lib = load(); _req(x, torch.float32, "x")
    NB, C, H, W = x.shape
    x = x.contiguous()
    y = torch.empty((NB, H, W, Cpad), device=x.device, dtype=BF16)
    check(lib.fd_nchw_to_nhwc(ptr(x), ptr(y), c_int32(NB), c_int32(C), c_int32(H), c_int32(W), c_int32(Cpad),
                              stream_ptr()), "fd_nchw_to_nhwc")
    return y


def nhwc_to_nchw(x, NB, C, H, W):
    """x: [NB*H*W, ld] fp32 or bf16 (first C columns valid) -> [NB,C,H,W] fp32."""
    lib = load()
    y = torch.empty((NB, C, H, W), device=x.device, dtype=torch.float32)
    check(lib.fd_nhwc_to_nchw(ptr(x), c_int32(1 if x.dtype == torch.float32 else 0), c_int64(x.stride(0)),
                              ptr(y), c_int32(NB), c_int32(C), c_int32(H), c_int32(W), stream_ptr()),
          "fd_nhwc_to_nchw")
    return y


def upsample2x(x, NB, H, W, C):
    lib = load(); _req(x, BF16, "x")
    y = torch.empty((NB * 4 * H * W, C), device=x.device, dtype=BF16)
    check(lib.fd_upsample2x(ptr(x), ptr(y), c_int32(NB), c_int32(H), c_int32(W), c_int32(C), stream_ptr()),
          "fd_upsample2x")
    return y


def upsample2x_bwd(dy, NB, H, W, C):
    lib = load(); _req(dy, BF16, "dy")
    dx = torch.empty((NB * H * W, C), device=dy.device, dtype=BF16)
    check(lib.fd_upsample2x_bwd(ptr(dy), ptr(dx), c_int32(NB), c_int32(H), c_int32(W), c_int32(C),
                                stream_ptr()), "fd_upsample2x_bwd")
    return dx


def space_to_depth(x, NB, H, W, C):
    lib = load(); _req(x, BF16, "x")
    y = torch.empty((4 * NB * (H // 2) * (W // 2), C), device=x.device, dtype=BF16)
    check(lib.fd_space_to_depth(ptr(x), ptr(y), c_int32(NB), c_int32(H), c_int32(W), c_int32(C), stream_ptr()),
          "fd_space_to_depth")
    return y


def depth_to_space(x, NB, H, W, C):
    lib = load(); _req(x, BF16, "x")
    y = torch.empty((NB * H * W, C), device=x.device, dtype=BF16)
    check(lib.fd_depth_to_space(ptr(x), ptr(y), c_int32(NB), c_int32(H), c_int32(W), c_int32(C), stream_ptr()),
          "fd_depth_to_space")
    return y


def concat_channels(a, b):
    lib = load(); _req(a, BF16, "a"); _req(b, BF16, "b")
    rows = a.shape[0]
    assert a.is_contiguous() and b.is_contiguous() and b.shape[0] == rows
    y = torch.empty((rows, a.shape[1] + b.shape[1]), device=a.device, dtype=BF16)
    check(lib.fd_concat_channels(ptr(a), c_int32(a.shape[1]), ptr(b), c_int32(b.shape[1]), ptr(y), c_int64(rows),
                                 stream_ptr()), "fd_concat_channels")
    return y


def slice_channels(x, c0, C):
    lib = load(); _req(x, BF16, "x")
    rows, Ctot = x.shape
    assert x.is_contiguous()
    y = torch.empty((rows, C), device=x.device, dtype=BF16)
    check(lib.fd_slice_channels(ptr(x), c_int32(Ctot), c_int32(c0), c_int32(C), ptr(y), c_int64(rows), stream_ptr()),
          "fd_slice_channels")
    return y


def add(a, b):
    lib = load(); _req(a, BF16, "a"); _req(b, BF16, "b")
    assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    y = torch.empty_like(a)
    check(lib.fd_add(ptr(a), ptr(b), ptr(y), c_int64(a.numel()), stream_ptr()), "fd_add")
    return y


def transpose(x):
    """bf16 [rows, cols] -> [cols, rows]; the result's row stride is padded to a multiple of 8 elements so that
    it is a valid TMA operand for any `rows` (the returned view has the exact shape)."""
    lib = load(); _req(x, BF16, "x")
    rows, cols = x.shape
    assert x.is_contiguous()
    ld = (rows + 7) // 8 * 8
    buf = torch.empty((cols, ld), device=x.device, dtype=BF16)
    check(lib.fd_transpose(ptr(x), ptr(buf), c_int32(rows), c_int32(cols), c_int64(ld), stream_ptr()), "fd_transpose")
    return buf[:, :rows]


def cast_scale(x, scale=1.0):
    lib = load(); _req(x, torch.float32, "x")
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    check(lib.fd_cast_scale(ptr(x), ptr(y), c_int64(x.numel()), c_float(scale), stream_ptr()), "fd_cast_scale")
    return y


def silu_f32_to_bf16(x):
    lib = load(); _req(x, torch.float32, "x")
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    check(lib.fd_silu_f32_to_bf16(ptr(x), ptr(y), c_int64(x.numel()), stream_ptr()), "fd_silu_f32_to_bf16")
    return y


def timestep_embedding(t, dim):
    lib = load(); _req(t, torch.float32, "t")
    y = torch.empty((t.numel(), dim), device=t.device, dtype=BF16)
    check(lib.fd_timestep_embedding(ptr(t), ptr(y), c_int32(t.numel()), c_int32(dim), stream_ptr()),
          "fd_timestep_embedding")
    return y


def geglu_bwd(acc, dout):
    lib = load(); _req(acc, BF16, "acc"); _req(dout, BF16, "dout")
    M, N = acc.shape
    assert acc.is_contiguous() and dout.is_contiguous() and dout.shape == (M, N // 2)
    dacc = torch.empty_like(acc)
    check(lib.fd_geglu_bwd(ptr(acc), ptr(dout), ptr(dacc), c_int64(M), c_int32(N), stream_ptr()), "fd_geglu_bwd")
    return dacc


# ------------------------------------------------------------------ distillation-step kernels (fp32)
def step_add_noise(z, noise, sa, sg):
    lib = load(); _req(z, torch.float32, "z")
    B = z.shape[0]
    out = torch.empty_like(z)
    check(lib.fd_step_add_noise(ptr(z.contiguous()), ptr(noise.contiguous()), ptr(sa), ptr(sg), ptr(out),
                                c_int32(B), c_int64(z.numel() // B), stream_ptr()), "fd_step_add_noise")
    return out


def step_cfg_dpm(eps_c, eps_u, x, x0_prev, coef6):
    """in-place on x and x0_prev (fp32, contiguous)."""
    lib = load()
    arr = (c_float * 6)(*[float(c) for c in coef6])
    check(lib.fd_step_cfg_dpm(ptr(eps_c), ptr(eps_u), ptr(x), ptr(x0_prev), arr, c_int64(x.numel()),
                              stream_ptr()), "fd_step_cfg_dpm")
    return x


def step_student_output(x_t, eps, sa, sg, c_skip, c_out):
    lib = load()
    B = x_t.shape[0]
    out = torch.empty_like(x_t)
    check(lib.fd_step_student_output(ptr(x_t), ptr(eps), ptr(sa), ptr(sg), ptr(c_skip), ptr(c_out), ptr(out),
                                     c_int32(B), c_int64(x_t.numel() // B), stream_ptr()),
          "fd_step_student_output")
    return out
