"""GPU parity tests proper: the CUDA library (through the C ABI) against the reference's golden
vectors (tests/golden/*.npz, produced by the unmodified reference) -- bit-exact bar, because the
kernels reproduce the reference's FP64 operation order with FMA contraction disabled."""
import glob
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(util.GOLDEN, "*.npz")))
assert len(CASES) >= 29   # every committed reference fixture, all solvers / integrators / orders


@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_bitwise(cuda_lib, name):
    meta, z, ctx, ti = util.run_golden_case(cuda_lib, name)
    util.assert_matches_golden(meta, z, ctx, ti)
    assert ctx.launches() > 0
    ctx.close()


@pytest.mark.parametrize("block", [(32, 16, 16), (8, 8, 8), (16, 16, 8)])
def test_cuda_decomposition_invariance(cuda_lib, block):
    meta, z, ctx, ti = util.run_golden_case(cuda_lib, "mhd_linwave_vl2_plm_hlld", block=block)
    util.assert_matches_golden(meta, z, ctx, ti)
    ctx.close()


def test_cuda_equals_host_build_of_same_operators(cuda_lib, emul_lib):
    """Device FP64 (div/sqrt IEEE, --fmad=false) == host FP64 on every intermediate array."""
    name = "mhd_linwave_alfven_vl2"
    outs = []
    for lib in (cuda_lib, emul_lib):
        meta, z, ctx, ti = util.run_golden_case(lib, name)
        arrs = {}
        for a in ("W", "BCC", "FLUX_X1", "FLUX_X2", "FLUX_X3", "E_X1E", "E_X2E", "E_X3E",
                  "WGHT_X1F", "E3_X1F", "E1_X3F"):
            arrs[a] = ctx.download(a, 3)
        outs.append(arrs)
        ctx.close()
    for a in outs[0]:
        np.testing.assert_array_equal(outs[0][a], outs[1][a], err_msg=a)


def test_divb_machine_zero_and_conservation_64cubed(cuda_lib):
    """Size-independent properties at a larger mesh: div(B) stays at machine zero under CT,
    mass/energy conserved to round-off (periodic box)."""
    from athena_b200 import pyhost
    meta, z = util.load_golden("mhd_linwave_alfven_vl2")
    # tile the 32x16x16 initial state 2x4x4 -> 64^3 (periodic, so still a valid solution)
    U = np.tile(z["u0"], (1, 4, 4, 2))
    B1 = np.tile(z["b1_0"][:, :, :-1], (4, 4, 2))
    B1 = np.concatenate([B1, B1[:, :, :1]], axis=2)
    B2 = np.tile(z["b2_0"][:, :-1, :], (4, 4, 2))
    B2 = np.concatenate([B2, B2[:, :1, :]], axis=1)
    B3 = np.tile(z["b3_0"][:-1, :, :], (4, 4, 2))
    B3 = np.concatenate([B3, B3[:1, :, :]], axis=0)
    ctx = util.make_ctx_from_global(cuda_lib, (64, 64, 64), (32, 32, 32), (0, 0, 0), (6.0, 6.0, 6.0),
                                    2, True, meta["gamma"], U, (B1, B2, B3))
    ti = pyhost.TimeIntegrator(ctx, "vl2", 2, 0.3)
    ti.initialize()
    m0 = U[0].sum()
    e0 = U[4].sum()
    for _ in range(10):
        ti.step()
    Uf, B = util.gather_from_ctx(ctx)
    dx = 6.0 / 64
    divb = ((B[0][:, :, 1:] - B[0][:, :, :-1]) + (B[1][:, 1:, :] - B[1][:, :-1, :])
            + (B[2][1:, :, :] - B[2][:-1, :, :])) / dx
    bmag = np.sqrt((z["b1_0"] ** 2).max() + (z["b2_0"] ** 2).max() + (z["b3_0"] ** 2).max())
    assert np.abs(divb).max() * dx / bmag < 2e-13   # a few ulp(|B|) per CT update, 10 cycles
    assert abs(Uf[0].sum() - m0) / m0 < 1e-13
    assert abs(Uf[4].sum() - e0) / e0 < 1e-13
    ctx.close()


@pytest.mark.parametrize("name", CASES)
def test_cuda_fused_stage_matches_reference_bitwise(cuda_lib, name):
    meta, z, ctx, ti = util.run_golden_case(cuda_lib, name, fused=True)
    util.assert_matches_golden(meta, z, ctx, ti)
    ctx.close()


@pytest.mark.parametrize("block", [(32, 16, 16), (8, 8, 8)])
def test_cuda_fused_decomposition_invariance(cuda_lib, block):
    meta, z, ctx, ti = util.run_golden_case(cuda_lib, "mhd_linwave_vl2_plm_hlld", block=block,
                                            fused=True)
    util.assert_matches_golden(meta, z, ctx, ti)
    ctx.close()



# The reference is decomposition-invariant only up to rounding: block-boundary EMFs are averaged
# between neighbours whose face coordinates differ in the last bit.  Running the unmodified
# reference with both decompositions of these two fixtures gives 161 / 1915 entries that differ by
# <= 1.4e-15, and this library reproduces either run bit for bit when given the same blocks.
DECOMPOSITION_SENSITIVE = {"mhd_linwave_alfven_vl2", "mhd_linwave_rk3_plm_roe_slow"}


@pytest.mark.parametrize("name", CASES)
def test_cuda_fused_stage_wide_blocks_match_reference(cuda_lib, name):
    # a second decomposition of every fixture (nx1 = 32: full warps along x1, tiles that straddle rows)
    meta, _ = util.load_golden(name)
    if meta["mesh"][0] != 32:   # the odd-extent fixtures: one block per mesh instead
        block = tuple(meta["mesh"])
    else:
        block = (32, meta["mesh"][1] // 2, meta["mesh"][2])
    meta, z, ctx, ti = util.run_golden_case(cuda_lib, name, block=block, fused=True)
    if name in DECOMPOSITION_SENSITIVE:
        U, B = util.gather_from_ctx(ctx)
        assert ti.ncycle == meta["ncycle"] and ti.dt == meta["dt_final"]
        np.testing.assert_allclose(U, z["uf"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(B[0], z["b1_f"], rtol=0, atol=1e-14)
    else:
        util.assert_matches_golden(meta, z, ctx, ti)
    ctx.close()


def test_full_bench_size_properties_256cubed(cuda_lib):
    """BASELINE's full single-GPU size (MHD blast, 256^3 zones in 64^3 MeshBlocks -- the bench
    workload), checked through size-independent properties: the fused stage and the
    operator-per-task sequence give the same bits, div(B) stays at machine zero under CT, mass and
    total energy are conserved on the periodic mesh."""
    import ctypes
    from athena_b200 import problems, pyhost
    gamma = 1.666666666667

    def make():
        mesh = pyhost.UniformMesh((256,) * 3, (64,) * 3, (-1.0,) * 3, (1.0,) * 3, ng=2)
        ctx = pyhost.Context(mesh, True, gamma, lib=cuda_lib)
        for lb, g in enumerate(ctx.gids):
            ic = problems.blast_block(mesh, mesh.loc[g], True, gamma)
            ctx.upload("U", lb, ic["u"])
            ctx.upload("B_X1F", lb, ic["b1"])
            ctx.upload("B_X2F", lb, ic["b2"])
            ctx.upload("B_X3F", lb, ic["b3"])
        return mesh, ctx

    def sums(ctx):
        s = (ctypes.c_double * 11)()
        ctx.call("athk_history", s, 11)
        return np.array(s[:])

    out = {}
    for fused in (True, False):
        mesh, ctx = make()
        ti = pyhost.TimeIntegrator(ctx, "vl2", 2, 0.3, fused=fused)
        ti.initialize()
        h0 = sums(ctx)
        for _ in range(3):
            ti.step()
        h1 = sums(ctx)
        ng, dx = 2, 2.0 / 256
        divmax = 0.0
        for lb in (0, 21, 42, 63):   # a sample of MeshBlocks, the blast centre among them
            b1, b2, b3 = (ctx.download(n, lb) for n in ("B_X1F", "B_X2F", "B_X3F"))
            s = slice(ng, ng + 64)
            div = ((b1[s, s, ng + 1:ng + 65] - b1[s, s, ng:ng + 64]) + (b2[s, ng + 1:ng + 65, s] - b2[s, ng:ng + 64, s])
                   + (b3[ng + 1:ng + 65, s, s] - b3[ng:ng + 64, s, s]))
            divmax = max(divmax, np.abs(div).max())
        out[fused] = (h1, ti.dt, [ctx.download("U", lb) for lb in (0, 42)])
        assert divmax < 1e-13, divmax                       # |B| = 1: a few ulp per CT update
        assert abs(h1[0] - h0[0]) < 1e-13 * h0[0]           # mass
        assert abs(h1[7] - h0[7]) < 1e-13 * h0[7]           # total energy
        ctx.close()
    assert out[True][1] == out[False][1]
    np.testing.assert_array_equal(out[True][0], out[False][0])
    for a, b in zip(out[True][2], out[False][2]):
        np.testing.assert_array_equal(a, b)
