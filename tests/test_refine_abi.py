"""The refinement entry points of the C ABI (include/athk.h: athk_refine_*) called directly: argument
checking (host logic, harness library) and one restriction + coarse W(U) + coarse transfer against a
numpy restatement of MeshRefinement::RestrictCellCenteredValues (mesh_refinement.cpp:116-186) and
RestrictFieldX1 (:272-299).  The full exchange is pinned against the reference binary in
tests/test_host_binary.py (smr_* cases)."""
import ctypes as C

import numpy as np
import pytest

from athena_b200 import capi, pyhost


class BoxOp(C.Structure):   # == athk_box_op
    _fields_ = [("kind", C.c_int), ("dst", C.c_int), ("src", C.c_int), ("dir", C.c_int), ("lo", C.c_int * 3),
                ("hi", C.c_int * 3), ("so", C.c_int * 3), ("pad", C.c_int), ("peer", C.c_int)]


def box(kind, dst, src, dir_, lo, hi, so=(0, 0, 0), pad=0, peer=0):
    return BoxOp(kind, dst, src, dir_, (C.c_int * 3)(*lo), (C.c_int * 3)(*hi), (C.c_int * 3)(*so), pad, peer)


def set_ops(ctx, phase, ops):
    arr = (BoxOp * max(len(ops), 1))(*ops)
    return ctx.lib.athk_refine_set_ops(ctx.h, phase, C.cast(arr, C.c_void_p), len(ops))


def make(lib, mhd):
    mesh = pyhost.UniformMesh((16, 8, 8), (8, 8, 8), (0.0, 0.0, 0.0), (2.0, 1.0, 1.0), ng=2)
    return mesh, pyhost.Context(mesh, mhd, 1.4, lib=lib)


def refine_checks(lib):
    mesh, ctx = make(lib, True)
    h = ctx.h
    # nothing works before athk_refine_enable
    assert set_ops(ctx, 1, [box(1, 0, 0, 0, (2, 2, 2), (5, 5, 5))]) != 0
    assert b"athk_refine_enable" in lib.athk_last_error()
    ctx.call("athk_refine_enable")
    dp = C.POINTER(C.c_double)
    for lb, g in enumerate(ctx.gids):
        xv = [np.ascontiguousarray(mesh.coords(mesh.loc[g], d)[1]) for d in range(3)]      # x?v
        cxv = []
        for d in range(3):   # coarse twin: cells of twice the width, same number of ghost cells
            xf = np.asarray(mesh.coords(mesh.loc[g], d)[0])
            cf = np.concatenate([xf[2] - (xf[4] - xf[2]) * np.arange(2, 0, -1), xf[2:-2:2], xf[-3] + (xf[4] - xf[2]) * np.arange(1, 3)])
            cxv.append(np.ascontiguousarray(0.5 * (cf[1:] + cf[:-1])))
        ctx.call("athk_refine_set_block_coords", lb, *[a.ctypes.data_as(dp) for a in xv + cxv])
    # boxes that leave their array, unknown kinds, field boxes need their component, peers must be given
    assert set_ops(ctx, 1, [box(1, 0, 0, 0, (0, 0, 0), (8, 7, 7))]) != 0          # coarse array is 8 wide: 0..7
    assert set_ops(ctx, 1, [box(19, 0, 0, 0, (2, 2, 2), (3, 3, 3))]) != 0
    assert set_ops(ctx, 1, [box(10, 0, 0, 3, (2, 2, 2), (3, 3, 3))]) != 0
    assert set_ops(ctx, 1, [box(9, 0, 1, 0, (2, 2, 2), (3, 3, 3), peer=2)]) != 0
    assert b"athk_refine_set_peers" in lib.athk_last_error()
    assert set_ops(ctx, 200, []) != 0
    assert set_ops(ctx, 5, [box(5, 0, 0, 0, (0, 2, 2), (3, 3, 3))]) != 0           # prolongation reads one cell around
    # restriction of u and of the x1 field over the whole coarse active zone, coarse W(U), transfer to the host
    rng = np.random.default_rng(3)
    u = np.zeros(ctx.shape("U"))
    u[0] = 1.0 + rng.random(u[0].shape)
    u[1:4] = 0.2 * rng.standard_normal(u[1:4].shape)
    u[4] = 3.0 + rng.random(u[4].shape)
    b1 = 0.1 * rng.standard_normal(ctx.shape("B_X1F"))
    for lb in range(ctx.nb):
        ctx.upload("U", lb, u)
        ctx.upload("B_X1F", lb, b1)
        ctx.upload("B_X2F", lb, np.zeros(ctx.shape("B_X2F")))
        ctx.upload("B_X3F", lb, np.zeros(ctx.shape("B_X3F")))
    cs, ce = 2, 2 + 8 // 2 - 1
    assert set_ops(ctx, 1, [box(1, 1, 1, 0, (cs, cs, cs), (ce, ce, ce)),
                            box(10, 1, 1, 0, (cs, cs, cs), (ce + 1, ce, ce)),
                            box(10, 1, 1, 1, (cs, cs, cs), (ce, ce + 1, ce)),
                            box(10, 1, 1, 2, (cs, cs, cs), (ce, ce, ce + 1))]) == 0
    assert set_ops(ctx, 4, [box(4, 1, 1, 0, (cs, cs, cs), (ce, ce, ce))]) == 0
    ctx.call("athk_refine_run", 1, 0)
    ctx.call("athk_refine_run", 4, 0)
    cn = 8 // 2 + 4
    cw = np.zeros((5, cn, cn, cn))
    cb1 = np.zeros((cn, cn, cn + 1))
    ctx.call("athk_refine_coarse_transfer", 1, 0, cw.ctypes.data_as(dp), cb1.ctypes.data_as(dp), None, None)
    # numpy: equal volumes / areas, the reference's summation order
    a = u[:, 2:10, 2:10, 2:10]
    f = lambda k, j, i: a[:, k::2, j::2, i::2]
    cu = (((f(0, 0, 0) * 1 + f(0, 1, 0) * 1) + (f(0, 0, 1) * 1 + f(0, 1, 1) * 1)) +
          ((f(1, 0, 0) * 1 + f(1, 1, 0) * 1) + (f(1, 0, 1) * 1 + f(1, 1, 1) * 1))) / 8.0
    np.testing.assert_allclose(cw[0, cs:ce + 1, cs:ce + 1, cs:ce + 1], cu[0], rtol=1e-14)
    np.testing.assert_allclose(cw[1, cs:ce + 1, cs:ce + 1, cs:ce + 1], cu[1] / cu[0], rtol=1e-13, atol=1e-15)
    fb = b1[2:10, 2:10, 2:11]
    g = lambda k, j: fb[k::2, j::2, ::2]
    cb = (g(0, 0) + g(0, 1) + g(1, 0) + g(1, 1)) / 4.0
    np.testing.assert_allclose(cb1[cs:ce + 1, cs:ce + 1, cs:ce + 2], cb, rtol=1e-13, atol=1e-16)
    # ... and back: what the host writes is what the device holds
    cw2 = cw + 1.0
    ctx.call("athk_refine_coarse_transfer", 1, 1, cw2.ctypes.data_as(dp), None, None, None)
    cw3 = np.zeros_like(cw)
    ctx.call("athk_refine_coarse_transfer", 1, 0, cw3.ctypes.data_as(dp), None, None, None)
    np.testing.assert_array_equal(cw3, cw2)
    assert lib.athk_refine_coarse_transfer(h, 7, 0, cw.ctypes.data_as(dp), None, None, None) != 0
    ctx.close()
    # field boxes are refused by a hydro context
    mesh, hyd = make(lib, False)
    hyd.call("athk_refine_enable")
    assert set_ops(hyd, 8, [box(9, 0, 1, 0, (2, 2, 2), (3, 3, 3))]) != 0
    assert b"MHD" in lib.athk_last_error()
    hyd.close()


def test_refine_entry_points_host_logic(emul_lib):
    refine_checks(emul_lib)


@pytest.mark.gpu
def test_refine_entry_points_on_gpu(cuda_lib):
    refine_checks(cuda_lib)
