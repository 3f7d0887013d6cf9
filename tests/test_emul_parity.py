"""CPU-side check of the operator bodies' index logic (tests/emul host build of the SAME functors
the CUDA kernels run) against the reference's golden vectors.  This is debugging infrastructure
for a container without a GPU; the GPU parity tests proper are in test_gpu_parity.py."""
import pytest

import util

CASES = ["mhd_linwave_vl2_plm_hlld", "hydro_linwave_vl2_plm_hllc", "mhd_linwave_rk2_plm_hlld",
         "mhd_linwave_rk3_plm_hlld", "hydro_linwave_rk3_plm_hllc", "mhd_linwave_alfven_vl2",
         "hydro_linwave_rk3_ppm_hllc", "mhd_linwave_vl2_ppm_hlld",
         # Roe / HLLE / LLF, hydro and MHD
         "hydro_linwave_vl2_plm_roe", "hydro_linwave_rk2_plm_roe_supersonic",
         "hydro_linwave_vl2_plm_hlle", "hydro_linwave_vl2_plm_llf", "mhd_linwave_vl2_plm_roe",
         "mhd_linwave_rk3_plm_roe_slow", "mhd_linwave_vl2_plm_hlle", "mhd_linwave_vl2_plm_llf",
         # Roe with negative intermediate densities -> LLF fallback, and supersonic upwinding
         "hydro_shocktube_vl2_plm_roe_llf", "mhd_shocktube_vl2_plm_roe_llf",
         # passive scalars (KH, Lecoanet dye), incl. BASELINE configs[3]'s RK3+PPM
         "hydro_kh_rk3_ppm_hllc_scalar", "hydro_kh_vl2_plm_hllc_2scalars", "mhd_kh_rk2_plm_hlld_scalar",
         # outflow / reflecting physical boundaries
         "hydro_sod_vl2_plm_hllc_outflow_x1", "mhd_briowu_rk2_plm_hlld_outflow_x2",
         "mhd_blast_vl2_plm_hlld_reflecting_box", "hydro_blast_rk3_plm_hllc_mixed_bcs",
         # odd extents (11x9x13, 12x10x14 blocks): row padding, tiles / warps straddling rows
         "mhd_linwave_vl2_plm_hlld_oddmesh", "hydro_linwave_rk3_ppm_hllc_oddmesh",
         # circularly polarised Alfven wave; field loop in a field-free box (HLLD / CT with B == 0)
         "mhd_cpaw_rk3_plm_hlld", "mhd_fieldloop_vl2_plm_hlld",
         # task SRC_TERM: constant accelerations
         "mhd_blast_rk2_plm_hlld_gravity"]


@pytest.mark.parametrize("name", CASES)
def test_emul_matches_reference_bitwise(emul_lib, name):
    meta, z, ctx, ti = util.run_golden_case(emul_lib, name)
    util.assert_matches_golden(meta, z, ctx, ti)
    ctx.close()


@pytest.mark.parametrize("block", [(32, 16, 16), (8, 8, 8), (16, 16, 8)])
def test_emul_decomposition_invariance(emul_lib, block):
    # the reference result depends on the global mesh only
    meta, z, ctx, ti = util.run_golden_case(emul_lib, "mhd_linwave_vl2_plm_hlld", block=block)
    util.assert_matches_golden(meta, z, ctx, ti)
    ctx.close()


@pytest.mark.parametrize("name", CASES)
def test_emul_fused_stage_matches_reference_bitwise(emul_lib, name):
    # athk_stage(): fused sweeps + integrate, must give the same bits as the operator sequence
    meta, z, ctx, ti = util.run_golden_case(emul_lib, name, fused=True)
    util.assert_matches_golden(meta, z, ctx, ti)
    ctx.close()
