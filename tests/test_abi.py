"""The C-ABI library loads (no compute calls: no GPU needed) and exports every symbol that
include/athk.h declares; the ctypes prototype table covers exactly the same set."""
import ctypes
import os
import re

import pytest

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


def test_operator_argument_checks(emul_lib):
    """Error behaviour of the operator entry points (host logic, exercised on the test-only host
    build): wrong order / missing scalars / bad boundary flags return ATHK_ERR_* with a message,
    the way the reference's constructors raise ATHENA_ERROR (reconstruction.cpp:117-127)."""
    from athena_b200 import pyhost
    lib = emul_lib
    mesh = pyhost.UniformMesh((8, 8, 8), (8, 8, 8), (0.0,) * 3, (1.0,) * 3, ng=2)
    ctx = pyhost.Context(mesh, False, 1.4, lib=lib)
    h = ctx.h
    assert lib.athk_calculate_fluxes(h, 3, 0.1, 0) == -1 and b"nghost >= 3" in lib.athk_last_error()   # PPM needs 3 ghosts
    assert lib.athk_calculate_fluxes(h, 4, 0.1, 0) == -1
    assert lib.athk_calculate_scalar_fluxes(h, 2) == -1 and b"no passive scalars" in lib.athk_last_error()
    assert lib.athk_add_scalar_flux_divergence(h, 0.1, 0) == -1
    assert lib.athk_compute_corner_e(h) == -1 and b"not MHD" in lib.athk_last_error()
    bad = (ctypes.c_int * 6)(0, 0, 7, 0, 0, 0)
    assert lib.athk_set_boundary_flags(h, 0, bad) == -1 and b"unknown flag" in lib.athk_last_error()
    assert lib.athk_set_boundary_flags(h, 5, (ctypes.c_int * 6)()) == -1
    # a face flagged as a physical boundary must not have a neighbour in the table
    assert lib.athk_set_boundary_flags(h, 0, (ctypes.c_int * 6)(2, 0, 0, 0, 0, 0)) == 0
    assert lib.athk_apply_physical_boundaries(h) == -4 and b"has a neighbour" in lib.athk_last_error()
    sums = (ctypes.c_double * 4)()
    assert lib.athk_history(h, sums, 4) == -1 and b"need room" in lib.athk_last_error()
    ctx.close()


def block_floor_check(lib):
    """athk_set_block_floors / athk_add_source_terms through the C ABI: W(U) floors each block
    with its own values; constant accelerations change momentum and energy as
    hydro/srcterms/constant_acc.cpp does; bad arguments are refused."""
    import numpy as np
    from athena_b200 import pyhost
    mesh = pyhost.UniformMesh((8, 4, 4), (4, 4, 4), (0.0,) * 3, (1.0,) * 3, ng=2)
    ctx = pyhost.Context(mesh, False, 1.4, lib=lib, nscalars=1, grav_acc=(0.0, 0.5, 0.0))
    h = ctx.h
    assert lib.athk_set_block_floors(h, 2, 1e-3, 1e-3, 1e-3) == -1 and b"out of range" in lib.athk_last_error()
    assert lib.athk_set_block_floors(h, 1, -1.0, 1e-3, 1e-3) == -1 and b"must be positive" in lib.athk_last_error()
    ctx.set_block_floors(1, 1.0e-3, 2.0e-3, 4.0e-3)
    u = np.zeros(ctx.shape("U"))
    u[0] = 1.0e-30           # density far below any floor
    u[4] = 1.0e-30
    s = np.zeros(ctx.shape("S"))
    for lb in range(2):
        ctx.upload("U", lb, u)
        ctx.upload("S", lb, s)
    ctx.call("athk_cons2prim", 0)
    fmin = np.sqrt(1024 * np.float64(np.finfo(np.float32).tiny))     # the default, block 0
    for lb, (df, pf, sf) in enumerate(((fmin, fmin, fmin), (1.0e-3, 2.0e-3, 4.0e-3))):
        w, r = ctx.download("W", lb), ctx.download("R", lb)
        assert np.all(w[0] == df) and np.all(w[4] == pf)
        assert np.all(r[0] == (sf * df) / df)
    # SRC_TERM on block 1: u(IM2) += dt*rho*g2, u(IEN) += dt*rho*g2*v2 on the active zone only
    before = ctx.download("U", 1)
    w = ctx.download("W", 1)
    ctx.call("athk_add_source_terms", 0.25)
    after = ctx.download("U", 1)
    act = (slice(2, 6),) * 3
    src = 0.25 * w[0] * 0.5
    np.testing.assert_array_equal(after[2][act], (before[2] + src)[act])
    np.testing.assert_array_equal(after[4][act], (before[4] + src * w[2])[act])
    ghost = np.ones(before[2].shape, dtype=bool)
    ghost[act] = False
    np.testing.assert_array_equal(after[2][ghost], before[2][ghost])
    np.testing.assert_array_equal(after[1], before[1])
    ctx.close()


def test_block_floors_and_source_terms_host_logic(emul_lib):
    block_floor_check(emul_lib)


@pytest.mark.gpu
def test_block_floors_and_source_terms_on_gpu(cuda_lib):
    block_floor_check(cuda_lib)
