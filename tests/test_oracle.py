"""Pins the oracle: (1) the numpy restatement (oracle/athena_np.py) must reproduce the golden
vectors -- outputs of the UNMODIFIED reference (tests/golden/make_golden.py) -- bit for bit;
(2) where the reference binary is available (this container, and the GPU box: oracle/_ref
travels), a fresh run of it must reproduce the committed fixtures."""
import glob
import shutil

import numpy as np
import pytest

import util
from oracle import athena_np as A
from oracle import refrun

CASES = ["mhd_linwave_vl2_plm_hlld", "hydro_linwave_vl2_plm_hllc", "mhd_linwave_rk2_plm_hlld",
         "mhd_linwave_rk3_plm_hlld", "hydro_linwave_rk3_plm_hllc", "mhd_linwave_alfven_vl2",
         # PPM (nghost = 3) and passive scalars
         "hydro_linwave_rk3_ppm_hllc", "mhd_linwave_vl2_ppm_hlld", "hydro_kh_rk3_ppm_hllc_scalar",
         "hydro_kh_vl2_plm_hllc_2scalars", "mhd_kh_rk2_plm_hlld_scalar",
         # alternative Riemann solvers (Roe incl. its supersonic overrides and LLF fallback)
         "hydro_linwave_vl2_plm_roe", "hydro_linwave_rk2_plm_roe_supersonic", "hydro_linwave_vl2_plm_hlle",
         "hydro_linwave_vl2_plm_llf", "mhd_linwave_vl2_plm_hlle", "mhd_linwave_vl2_plm_llf",
         "hydro_shocktube_vl2_plm_roe_llf", "mhd_linwave_vl2_plm_roe", "mhd_linwave_rk3_plm_roe_slow",
         "mhd_shocktube_vl2_plm_roe_llf",
         # reflecting / outflow physical boundaries
         "hydro_sod_vl2_plm_hllc_outflow_x1", "mhd_briowu_rk2_plm_hlld_outflow_x2",
         "hydro_blast_rk3_plm_hllc_mixed_bcs", "mhd_blast_vl2_plm_hlld_reflecting_box",
         "mhd_linwave_vl2_plm_hlld_oddmesh", "hydro_linwave_rk3_ppm_hllc_oddmesh",
         # circularly polarised Alfven wave; field loop in a field-free box (HLLD / CT with B == 0)
         "mhd_cpaw_rk3_plm_hlld", "mhd_fieldloop_vl2_plm_hlld",
         # task SRC_TERM: constant accelerations
         "mhd_blast_rk2_plm_hlld_gravity"]


@pytest.mark.parametrize("name", CASES)
def test_numpy_oracle_matches_reference_bitwise(name):
    meta, z = util.load_golden(name)
    mesh = A.Mesh(meta["mesh"], meta["xmin"], meta["xmax"], ng=meta["ng"], bcs=meta.get("bcs"))
    b = [z["b1_0"], z["b2_0"], z["b3_0"]] if meta["mhd"] else None
    st = A.State(mesh, meta["gamma"], z["u0"], b, s=z["s0"] if meta.get("nscalars") else None,
                 flux=meta.get("flux"))
    it = A.Integrator(st, meta["integrator"], meta["xorder"], meta["cfl"], block_nx=meta["block"],
                      grav_acc=meta.get("grav_acc"))
    it.run(meta["tlim"])
    assert it.ncycle == meta["ncycle"]
    np.testing.assert_array_equal(np.array(it.dt_history), z["hst"][:-1, 1])
    assert it.dt == meta["dt_final"]
    np.testing.assert_array_equal(st.interior_u(), z["uf"])
    if meta.get("nscalars"):
        np.testing.assert_array_equal(st.interior_s(), z["sf"])
    if meta["mhd"]:
        np.testing.assert_array_equal(st.interior_b(0), z["b1_f"])
        np.testing.assert_array_equal(st.interior_b(1), z["b2_f"])
        np.testing.assert_array_equal(st.interior_b(2), z["b3_f"])


def test_golden_fixture_reproduced_by_reference_binary():
    cfg = "linwave_mhd_hlld"
    if not refrun.have_ref(cfg):
        pytest.skip("reference binary oracle/_ref/athena_%s not built" % cfg)
    meta, z = util.load_golden("mhd_linwave_vl2_plm_hlld")
    blocks = refrun.linwave_blocks(True, mesh=meta["mesh"], block=meta["block"], tlim=meta["tlim"],
                                   integrator=meta["integrator"], xorder=meta["xorder"])
    wd, out = refrun.run_reference(cfg, blocks)
    rf = refrun.assemble(refrun.read_rst(glob.glob(wd + "/*.final.rst")[0]))
    shutil.rmtree(wd, ignore_errors=True)
    np.testing.assert_array_equal(rf["u"], z["uf"])
    np.testing.assert_array_equal(rf["b1"], z["b1_f"])


def test_oracle_operator_level_hlld_vs_host_build(emul_lib):
    """Operator-level cross-check of two independent implementations (numpy restatement vs the
    C++ operator bodies): x1 fluxes, face EMFs and CT weights of one CalculateFluxes call."""
    from athena_b200 import pyhost
    meta, z = util.load_golden("mhd_linwave_alfven_vl2")
    B = (z["b1_0"], z["b2_0"], z["b3_0"])
    ctx = util.make_ctx_from_global(emul_lib, meta["mesh"], meta["mesh"], meta["xmin"], meta["xmax"],
                                    meta["ng"], True, meta["gamma"], z["u0"], B)
    ti = pyhost.TimeIntegrator(ctx, "vl2", 2, meta["cfl"])
    ti.initialize(tlim=meta["tlim"])
    ctx.call("athk_calculate_fluxes", 2, ti.dt, 0)
    ctx.call("athk_compute_corner_e")
    mesh = A.Mesh(meta["mesh"], meta["xmin"], meta["xmax"])
    st = A.State(mesh, meta["gamma"], z["u0"], list(B))
    with np.errstate(all="ignore"):
        fx = A.fluxes(st, 2, ti.dt)
        e = A.corner_e(st, fx)
    ng = meta["ng"]
    n1, n2, n3 = meta["mesh"]
    # active faces / edges incl. the upper boundary ones
    sl = (slice(ng, ng + n3), slice(ng, ng + n2), slice(ng, ng + n1))
    slx = (slice(ng, ng + n3), slice(ng, ng + n2), slice(ng, ng + n1 + 1))
    f1 = ctx.download("FLUX_X1", 0)
    for n in range(5):
        np.testing.assert_array_equal(f1[n][slx], fx[0][0][n][slx])
    np.testing.assert_array_equal(ctx.download("E3_X1F", 0)[slx], fx[0][1][slx])
    np.testing.assert_array_equal(ctx.download("WGHT_X2F", 0)[sl], fx[1][3][sl])
    np.testing.assert_array_equal(ctx.download("E1_X3F", 0)[sl], fx[2][2][sl])
    sle = (slice(ng, ng + n3 + 1), slice(ng, ng + n2 + 1), slice(ng, ng + n1))
    np.testing.assert_array_equal(ctx.download("E_X1E", 0)[sle], e[0][sle])
    sle3 = (slice(ng, ng + n3), slice(ng, ng + n2 + 1), slice(ng, ng + n1 + 1))
    np.testing.assert_array_equal(ctx.download("E_X3E", 0)[sle3], e[2][sle3])
    ctx.close()
