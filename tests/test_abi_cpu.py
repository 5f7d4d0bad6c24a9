"""CPU checks: the C-ABI libraries load, export every declared symbol, and refuse to run without a GPU."""
import ctypes as C
import os
import re

import pytest

from is3d2_b200 import capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header: str):
    text = open(os.path.join(REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(is3d_[a-z0-9_]+)\s*\(", text)))


def test_abi_exports_every_declared_symbol(libs):
    lib, host = libs
    for sym in _declared("is3d_b200.h"):
        assert hasattr(lib, sym), f"libis3d_b200.so does not export {sym}"
    assert sorted(capi.ABI_SYMBOLS) == _declared("is3d_b200.h")
    for sym in _declared("is3d_host.h"):
        assert hasattr(host, sym), f"libis3d_host.so does not export {sym}"
    assert sorted(capi.HOST_SYMBOLS) == _declared("is3d_host.h")


def test_params_struct_layout_matches_header():
    """ctypes mirror has the same field names, in order, as is3d_params / is3d_stats in the header."""
    text = open(os.path.join(REPO, "include", "is3d_b200.h")).read()
    for struct, cls in (("is3d_params", capi.Params), ("is3d_stats", capi.Stats), ("is3d_particle", capi.Particle)):
        body = re.search(r"typedef struct \{([^}]*)\} " + struct + ";", text).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(int64_t|int32_t|int|double)\s+", "", decl)
            names += [n.strip() for n in decl.split(",")]
        assert names == [f for f, _ in cls._fields_], struct


def test_create_fails_loudly_without_gpu(libs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib, _ = libs
    p = capi.Params()
    lib.is3d_default_params(C.byref(p))
    ctx = C.c_void_p()
    st = lib.is3d_create(C.byref(p), C.byref(ctx))
    assert st != 0 and not ctx.value
    assert b"no CPU path" in lib.is3d_last_error(None) or b"sm_" in lib.is3d_last_error(None)
