"""Raw (non-autograd) Python entry points over the C ABI: tensors in, tensors out.

Every function launches hand-written sm_100a kernels from libflashb200.so on the current CUDA
stream.  Tensors must be CUDA tensors; bf16 activations are channels-last ([rows, C] / NHWC).
"""
import ctypes
from ctypes import byref, c_float, c_int32, c_int64, c_void_p

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


def gemm(a1, b1, *, a2=None, b2=None, bias=None, rowvec=None, rows_per_group=0, geglu=False,
         residual=None, out=None, out_fp32=False, conv=None, M=None, force_bn=0):
    """acc = a1 @ b1.T (+ a2 @ b2.T) with the fused epilogue of fd_gemm (include/flashb200.h).

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
        assert rowvec.is_contiguous() and rowvec.shape[-1] == N
        args.rowvec, args.rows_per_group = ptr(rowvec), rows_per_group
    args.geglu = 1 if geglu else 0
    if residual is not None:
        assert residual.shape == (M, n_out) and residual.stride(1) == 1
        args.residual, args.ldr = ptr(residual), residual.stride(0)
    args.out, args.ldo, args.out_fp32 = ptr(out), out.stride(0), 1 if out_fp32 else 0
    args.force_bn = force_bn
    check(lib.fd_gemm(byref(args), stream_ptr()), "fd_gemm")
    return out
