"""User-enrolled boundary functions through the C ABI (athk_enroll_user_boundary: BValFunc,
athena.hpp:214-217; Mesh::EnrollUserBoundaryFunction, mesh.cpp:1288; dispatch bvals.cpp:673).

The enrolled function here does, on the arrays the library hands it, what the reference's outflow
functions do (bvals/cc/outflow_cc.cpp, bvals/fc/outflow_fc.cpp) -- so a run whose outflow faces are
re-flagged `user` must end on the bits of the reference's outflow fixtures.  That pins the hook's
place in the x1, x2, x3 order, the transverse limits it is given and the U(W) pass that follows it.
On the CPU the "device" arrays of the test harness library are host memory (numpy views); on the GPU
they are wrapped as torch tensors and written on the library's stream: a device-resident boundary
function."""
import numpy as np
import pytest

import util
from athena_b200 import pyhost

CASES = ["hydro_sod_vl2_plm_hllc_outflow_x1", "mhd_briowu_rk2_plm_hlld_outflow_x2", "hydro_blast_rk3_plm_hllc_mixed_bcs"]


class _DevPtr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _view(ptr, nvar, a, on_gpu, P3):
    """(nvar, k, j, i) strided view of one padded-box array of the block"""
    n = int(nvar * a.stride_n)
    shape = (nvar, P3, int(a.stride_k // a.stride_j), int(a.stride_j))
    if on_gpu:
        import torch
        flat = torch.as_tensor(_DevPtr(ptr, n), device="cuda")
        return flat.as_strided(shape, (int(a.stride_n), int(a.stride_k), int(a.stride_j), 1))
    import ctypes
    flat = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(n,))
    return np.lib.stride_tricks.as_strided(flat, shape, tuple(8 * s for s in (a.stride_n, a.stride_k, a.stride_j, 1)))


def _outflow(a, mhd, on_gpu):
    d, side = a.face // 2, a.face % 2
    lo = [a.il, a.jl, a.kl]
    hi = [a.iu, a.ju, a.ku]
    P3 = int(a.stride_n // a.stride_k)

    def fill(arr, comp_dir):
        # comp_dir: -1 cell-centred; 0..2 the face field's own direction
        for m in range(1, a.ngh + 1):
            if comp_dir == d:      # normal component: faces lo-m <- lo, hi+1+m <- hi+1
                g, s = (hi[d] + 1 + m, hi[d] + 1) if side else (lo[d] - m, lo[d])
            else:
                g, s = (hi[d] + m, hi[d]) if side else (lo[d] - m, lo[d])
            sl_g, sl_s = [slice(None)] * 4, [slice(None)] * 4
            for t in range(3):
                ax = 3 - t         # array axes are (n, k, j, i)
                if t == d:
                    sl_g[ax], sl_s[ax] = slice(g, g + 1), slice(s, s + 1)
                else:
                    ext = 1 if comp_dir == t else 0
                    sl_g[ax] = sl_s[ax] = slice(lo[t], hi[t] + 1 + ext)
            arr[tuple(sl_g)] = arr[tuple(sl_s)]

    fill(_view(a.w, 5, a, on_gpu, P3), -1)
    if mhd:
        for comp, ptr in enumerate((a.b1, a.b2, a.b3)):
            fill(_view(ptr, 1, a, on_gpu, P3), comp)
    return 0


def _run(lib, name, on_gpu, fused):
    meta, z = util.load_golden(name)
    mhd = meta["mhd"]
    bcs = ["user" if b == "outflow" else b for b in meta["bcs"]]
    assert "user" in bcs
    B = (z["b1_0"], z["b2_0"], z["b3_0"]) if mhd else (None, None, None)
    ctx = util.make_ctx_from_global(lib, meta["mesh"], meta["block"], meta["xmin"], meta["xmax"], meta["ng"], mhd,
                                    meta["gamma"], z["u0"], B, flux=meta.get("flux"), bcs=bcs)
    calls = []
    if on_gpu:
        import torch
        stream = torch.cuda.ExternalStream(lib.athk_stream(ctx.h))

        def fn(a):
            calls.append((a.lb, a.face))
            assert a.stream == stream.cuda_stream
            with torch.cuda.stream(stream):
                return _outflow(a, mhd, True)
    else:
        def fn(a):
            calls.append((a.lb, a.face))
            return _outflow(a, mhd, False)
    for face, b in enumerate(bcs):
        if b == "user":
            ctx.enroll_user_boundary(face, fn)
    ti = pyhost.TimeIntegrator(ctx, meta["integrator"], meta["xorder"], meta["cfl"], fused=fused)
    ti.initialize(tlim=meta["tlim"])
    ti.run(meta["tlim"])
    util.assert_matches_golden(meta, z, ctx, ti)
    # every flagged (block, face) pair was served each time the boundaries were applied
    flagged = sum(1 for g in ctx.gids for f in ctx.mesh.block_bcs(ctx.mesh.loc[g]) if f == 3)
    nstage = {"vl2": 2, "rk2": 2, "rk3": 3}[meta["integrator"]]
    assert len(calls) == flagged * (1 + nstage * meta["ncycle"])
    ctx.close()


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_user_boundary_function_reproduces_outflow_fixture(emul_lib, name, fused):
    _run(emul_lib, name, False, fused)


def test_user_flag_without_enrolled_function_is_refused(emul_lib):
    meta, z = util.load_golden(CASES[0])
    bcs = ["user" if b == "outflow" else b for b in meta["bcs"]]
    ctx = util.make_ctx_from_global(emul_lib, meta["mesh"], meta["block"], meta["xmin"], meta["xmax"], meta["ng"],
                                    False, meta["gamma"], z["u0"], bcs=bcs)
    ti = pyhost.TimeIntegrator(ctx, meta["integrator"], meta["xorder"], meta["cfl"])
    with pytest.raises(Exception, match="boundary function is not enrolled"):
        ti.initialize(tlim=meta["tlim"])
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES[:2])
def test_device_resident_user_boundary_function_on_gpu(cuda_lib, name):
    _run(cuda_lib, name, True, True)
