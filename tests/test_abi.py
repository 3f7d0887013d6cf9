"""The C-ABI library loads (no compute calls: no GPU needed) and exports every symbol that
include/athk.h declares; the ctypes prototype table covers exactly the same set."""
import ctypes
import os
import re

from athena_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "athk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(athk_[a-z0-9_]+)\s*\(", src)))


def test_header_and_prototype_table_agree():
    assert declared_symbols() == sorted(capi.PROTOTYPES)


def test_cuda_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "build libathk.so first (__graft_entry__.build())"
    lib = ctypes.CDLL(capi.LIB_PATH)     # loading needs libcudart only, not a device
    for name in declared_symbols():
        assert hasattr(lib, name), "libathk.so does not export " + name
    capi.bind(lib)


def test_no_device_is_an_error_not_a_fallback():
    """Without a CUDA device athk_create must fail loudly (ATHK_ERR_NODEVICE), never compute."""
    import torch
    if torch.cuda.is_available():
        return
    lib = capi.bind(ctypes.CDLL(capi.LIB_PATH))
    cfg = capi.Config(nx1=8, nx2=8, nx3=8, nghost=2, nblocks=1, mhd=1, nscalars=0,
                      rsolver=capi.RS_HLLD, device=0, gamma=1.4)
    h = ctypes.c_void_p()
    rc = lib.athk_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -3 and not h.value
    assert b"no CUDA device" in lib.athk_last_error()


def test_bad_arguments_are_rejected(emul_lib):
    lib = emul_lib
    h = ctypes.c_void_p()
    for kw in (dict(nx1=2), dict(nghost=1), dict(nblocks=0), dict(gamma=0.9), dict(rsolver=capi.RS_HLLC),
               dict(rsolver=7), dict(mhd=0)):   # hllc is hydro-only, hlld MHD-only (configure.py:376-389)
        base = dict(nx1=8, nx2=8, nx3=8, nghost=2, nblocks=1, mhd=1, nscalars=0,
                    rsolver=capi.RS_HLLD, device=0, gamma=1.4)
        base.update(kw)
        cfg = capi.Config(**base)
        assert lib.athk_create(ctypes.byref(cfg), ctypes.byref(h)) == -1, kw
        assert lib.athk_last_error()
