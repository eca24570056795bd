"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/flashb200.h declares; the product path fails loudly without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flashb200.h")
LIB = os.path.join(ROOT, "flash-diffusion_b200", "lib", "libflashb200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_library_loads_and_reports_without_gpu(lib):
    lib.fd_version.restype = ctypes.c_int
    lib.fd_last_error.restype = ctypes.c_char_p
    assert lib.fd_version() >= 100
    assert isinstance(lib.fd_last_error(), bytes)
    if not torch.cuda.is_available():
        assert lib.fd_sm_arch() < 0


def test_ctypes_struct_layout_matches_header():
    """FdGemmArgs field order in the ctypes mirror == the C declaration."""
    from flash.b200.lib import FdAttnArgs, FdAttnBwdArgs, FdGemmArgs
    text = open(HEADER).read()
    body = text[text.index("typedef struct {", text.index("fd_gemm —")):text.index("} FdGemmArgs;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = [re.sub(r"\[.*\]", "", f.strip().split()[-1].lstrip("*")) for decl in body.split(";")
                for f in decl.replace("typedef struct {", "").split(",") if f.strip()]
    py_fields = [f[0] for f in FdGemmArgs._fields_]
    assert c_fields == py_fields, (c_fields, py_fields)
    assert ctypes.sizeof(FdAttnBwdArgs) > ctypes.sizeof(FdAttnArgs)


def test_product_denoiser_refuses_cpu():
    from flash.models.unets import DiffusersUNet2DCondWrapper
    from flash.recipes import TINY_UNET_KWARGS
    net = DiffusersUNet2DCondWrapper(**TINY_UNET_KWARGS)
    x = torch.randn(1, 4, 32, 32)
    cond = {"cond": {"crossattn": torch.randn(1, 77, 96), "vector": torch.randn(1, 96)}}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, torch.tensor([10.0]), cond)


def test_missing_library_fails_loudly(monkeypatch):
    from flash.b200 import lib as fdlib
    monkeypatch.setattr(fdlib, "_lib", None)
    monkeypatch.setattr(fdlib, "LIB_PATH", "/nonexistent/libflashb200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fdlib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "flash-diffusion_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
