"""ctypes binding of the C ABI declared in include/flashb200.h.

The product path has NO fallback: if libflashb200.so is missing or a call fails, a
RuntimeError is raised (the oracle under /oracle is test infrastructure and is never imported
from here).
"""
import ctypes
import os
from ctypes import Structure, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

FD_MAX_TAPS = 16

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.environ.get(
    "FLASHB200_LIB", os.path.join(_PKG_DIR, "lib", "libflashb200.so")
)


class FdGemmArgs(Structure):
    _fields_ = [
        ("M", c_int32), ("N", c_int32),
        ("a1", c_void_p), ("lda1", c_int64), ("b1", c_void_p), ("ldb1", c_int64), ("K1", c_int32),
        ("a2", c_void_p), ("lda2", c_int64), ("b2", c_void_p), ("ldb2", c_int64), ("K2", c_int32),
        ("conv_taps", c_int32), ("NB_in", c_int32), ("H", c_int32), ("W", c_int32), ("C", c_int32),
        ("tap_dn", c_int32 * FD_MAX_TAPS), ("tap_dh", c_int32 * FD_MAX_TAPS),
        ("tap_dw", c_int32 * FD_MAX_TAPS),
        ("bias", c_void_p),
        ("rowvec", c_void_p), ("rows_per_group", c_int32), ("ldrv", c_int64),
        ("geglu", c_int32),
        ("residual", c_void_p), ("ldr", c_int64),
        ("out", c_void_p), ("ldo", c_int64), ("out_fp32", c_int32),
        ("force_bn", c_int32),
        ("ln_stats", c_void_p), ("ln_colsum", c_void_p), ("ln_inv_c", c_float), ("ln_eps", c_float),
        ("rowstats_out", c_void_p), ("rowstats_prezeroed", c_int32),
        ("act", c_int32),
        ("rowscale", c_void_p), ("rows_per_group_scale", c_int32), ("ldrs", c_int64),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("colstats_out", c_void_p), ("colstats_rows", c_int32),
    ]


class FdAttnArgs(Structure):
    _fields_ = [
        ("q", c_void_p), ("ldq", c_int64), ("q_batch_stride", c_int64),
        ("k", c_void_p), ("ldk", c_int64), ("k_batch_stride", c_int64),
        ("v", c_void_p), ("ldv", c_int64), ("v_batch_stride", c_int64),
        ("o", c_void_p), ("ldo", c_int64), ("o_batch_stride", c_int64),
        ("lse", c_void_p),
        ("B", c_int32), ("H", c_int32), ("Nq", c_int32), ("Nkv", c_int32),
        ("scale", c_float),
    ]


class FdAttnBwdArgs(Structure):
    _fields_ = [
        ("f", FdAttnArgs),
        ("d_o", c_void_p), ("lddo", c_int64), ("do_batch_stride", c_int64),
        ("dq", c_void_p), ("lddq", c_int64), ("dq_batch_stride", c_int64),
        ("dk", c_void_p), ("lddk", c_int64), ("dk_batch_stride", c_int64),
        ("dv", c_void_p), ("lddv", c_int64), ("dv_batch_stride", c_int64),
        ("delta", c_void_p),
        ("dq_accum", c_void_p),
    ]


_lib = None


def load():
    """Load libflashb200.so (once).  Raises RuntimeError when it is absent — no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libflashb200.so not found at {LIB_PATH}: build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "The B200 backend has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.fd_last_error.restype = c_char_p
    lib.fd_version.restype = c_int32
    lib.fd_sm_arch.restype = c_int32
    lib.fd_gemm_workspace_bytes.restype = ctypes.c_size_t
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().fd_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return c_void_p(t.data_ptr())
