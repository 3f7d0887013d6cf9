"""The fast-arithmetic kernel set (athk_config.arith = ATHK_ARITH_FAST: FMA contraction on, divisions
and square roots refined to ~1 ulp instead of correctly rounded) against the reference's golden
vectors.  The strict set is bit-exact (tests/test_gpu_parity.py); this set is held to tolerances:
fields to 1e-12 of their scale, the dt sequence to 1e-12 relative, conserved histories to 1e-10
(north star), div(B) at machine zero, and the linear-wave L1 error vector to what rounding-level
field differences allow (its components are ~1e-8 of the field scale, so 1e-12 of the FIELD scale
is ~1e-4 of an error component; the measured figure is printed and bounded)."""
import ctypes

import numpy as np
import pytest

import util
from athena_b200 import capi

CASES = ["mhd_linwave_vl2_plm_hlld", "hydro_linwave_vl2_plm_hllc", "mhd_blast_rk2_plm_hlld_gravity",
         "mhd_linwave_vl2_plm_hlld_oddmesh", "mhd_fieldloop_vl2_plm_hlld", "mhd_cpaw_rk3_plm_hlld"]


def _existing(cases):
    import os
    return [c for c in cases if os.path.exists(os.path.join(util.GOLDEN, c + ".npz"))]


@pytest.mark.gpu
@pytest.mark.parametrize("name", _existing(CASES))
def test_fast_arithmetic_set_agrees_with_reference_to_rounding(cuda_lib, name):
    meta, z, ctx, ti = util.run_golden_case(cuda_lib, name, fused=True, arith=capi.ARITH_FAST)
    try:
        U, B = util.gather_from_ctx(ctx)
        assert ti.ncycle == meta["ncycle"]
        hst = z["hst"]
        dts = np.array(ti.dt_history)
        np.testing.assert_allclose(dts, hst[:-1, 1][:len(dts)], rtol=1e-12, atol=0.0)
        scale = np.abs(z["uf"]).max(axis=(1, 2, 3), keepdims=True)
        scale = np.maximum(scale, np.abs(z["uf"]).max())   # velocities ~0: use the state's scale
        du = np.abs(U - z["uf"]).max() / scale.max()
        assert np.all(np.abs(U - z["uf"]) <= 1e-12 * scale), du
        if meta["mhd"]:
            bs = max(np.abs(z["b1_f"]).max(), np.abs(z["b2_f"]).max(), np.abs(z["b3_f"]).max(), 1e-300)
            for a, b in zip(B, (z["b1_f"], z["b2_f"], z["b3_f"])):
                assert np.abs(a - b).max() <= 1e-12 * bs
            # div(B) by CT stays at machine zero whatever the EMFs' last bits are
            dx = (np.array(meta["xmax"]) - np.array(meta["xmin"])) / np.array(meta["mesh"])
            div = (B[0][:, :, 1:] - B[0][:, :, :-1]) / dx[0] + (B[1][:, 1:, :] - B[1][:, :-1, :]) / dx[1] + \
                  (B[2][1:, :, :] - B[2][:-1, :, :]) / dx[2]
            assert np.abs(div).max() * dx.min() / bs < 1e-13
        nv = 8 + (3 if meta["mhd"] else 0)
        sums = (ctypes.c_double * nv)()
        ctx.call("athk_history", sums, nv)
        hs = np.abs(hst[:, 2:2 + nv]).max(axis=0) + 1e-300
        vol = np.prod((np.array(meta["xmax"]) - np.array(meta["xmin"])) / np.array(meta["mesh"]))
        hs[1:4] = vol * np.abs(z["uf"][1:4]).sum(axis=(1, 2, 3)) + 1e-300
        assert np.all(np.abs(np.array(sums[:]) - hst[-1, 2:2 + nv]) <= 1e-10 * hs)
        print("fast arithmetic %s: max field deviation %.2e of scale" % (name, du))
    finally:
        ctx.close()


@pytest.mark.gpu
def test_fast_arithmetic_linear_wave_error_vector(cuda_lib):
    """How far the fast set can move an L1 error vector of the linear-wave test (linear_wave.cpp:
    224-343: mean |u - u_exact| per variable): bounded by mean |u_fast - u_ref| per variable, put in
    proportion to the wave itself, mean |u_ref - u0|.  The strict set reproduces the reference's
    state exactly (deviation 0)."""
    name = "mhd_linwave_vl2_plm_hlld"
    dev = {}
    for arith in (capi.ARITH_STRICT, capi.ARITH_FAST):
        meta, z, ctx, ti = util.run_golden_case(cuda_lib, name, fused=True, arith=arith)
        U, _ = util.gather_from_ctx(ctx)
        ctx.close()
        dev[arith] = np.abs(U - z["uf"]).mean(axis=(1, 2, 3))
    wave = np.abs(z["uf"] - z["u0"]).mean(axis=(1, 2, 3)).max()
    assert dev[capi.ARITH_STRICT].max() == 0.0
    rel = dev[capi.ARITH_FAST] / wave
    print("fast arithmetic: L1 deviation per variable relative to the wave's L1 change", rel)
    assert rel.max() < 1e-7
