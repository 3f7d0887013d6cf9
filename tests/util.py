"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

from athena_b200 import pyhost

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def split_to_blocks(mesh, U, B1=None, B2=None, B3=None):
    """Global interior arrays -> per-block arrays with (zeroed) ghost zones, AthenaArray layout."""
    ng = mesh.ng
    nx1, nx2, nx3 = mesh.block_nx
    out = []
    for (l1, l2, l3) in mesh.loc:
        o1, o2, o3 = l1 * nx1, l2 * nx2, l3 * nx3
        nv = U.shape[0]
        u = np.zeros((nv, nx3 + 2 * ng, nx2 + 2 * ng, nx1 + 2 * ng))
        u[:, ng:ng + nx3, ng:ng + nx2, ng:ng + nx1] = U[:, o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1]
        blk = dict(u=u)
        if B1 is not None:
            b1 = np.zeros((nx3 + 2 * ng, nx2 + 2 * ng, nx1 + 2 * ng + 1))
            b2 = np.zeros((nx3 + 2 * ng, nx2 + 2 * ng + 1, nx1 + 2 * ng))
            b3 = np.zeros((nx3 + 2 * ng + 1, nx2 + 2 * ng, nx1 + 2 * ng))
            b1[ng:ng + nx3, ng:ng + nx2, ng:ng + nx1 + 1] = B1[o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1 + 1]
            b2[ng:ng + nx3, ng:ng + nx2 + 1, ng:ng + nx1] = B2[o3:o3 + nx3, o2:o2 + nx2 + 1, o1:o1 + nx1]
            b3[ng:ng + nx3 + 1, ng:ng + nx2, ng:ng + nx1] = B3[o3:o3 + nx3 + 1, o2:o2 + nx2, o1:o1 + nx1]
            blk.update(b1=b1, b2=b2, b3=b3)
        out.append(blk)
    return out


def gather_from_ctx(ctx, reg="U"):
    """Per-block device arrays -> global interior arrays (u, b1, b2, b3)."""
    mesh = ctx.mesh
    ng = mesh.ng
    nx1, nx2, nx3 = mesh.block_nx
    M1, M2, M3 = mesh.mesh_nx
    U = np.empty((ctx.nscalars if reg == "S" else 5, M3, M2, M1))
    B = [None, None, None]
    if ctx.mhd:
        B = [np.empty((M3, M2, M1 + 1)), np.empty((M3, M2 + 1, M1)), np.empty((M3 + 1, M2, M1))]
    for lb, g in enumerate(ctx.gids):
        l1, l2, l3 = mesh.loc[g]
        o1, o2, o3 = l1 * nx1, l2 * nx2, l3 * nx3
        u = ctx.download(reg, lb)
        U[:, o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1] = u[:, ng:ng + nx3, ng:ng + nx2, ng:ng + nx1]
        if ctx.mhd and reg != "S":
            b1 = ctx.download("B_X1F", lb)
            b2 = ctx.download("B_X2F", lb)
            b3 = ctx.download("B_X3F", lb)
            B[0][o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1 + 1] = b1[ng:ng + nx3, ng:ng + nx2, ng:ng + nx1 + 1]
            B[1][o3:o3 + nx3, o2:o2 + nx2 + 1, o1:o1 + nx1] = b2[ng:ng + nx3, ng:ng + nx2 + 1, ng:ng + nx1]
            B[2][o3:o3 + nx3 + 1, o2:o2 + nx2, o1:o1 + nx1] = b3[ng:ng + nx3 + 1, ng:ng + nx2, ng:ng + nx1]
    return U, B


RSOLVERS = {"hllc": 0, "hlld": 1, "roe": 2, "hlle": 3, "llf": 4}   # enum athk_rsolver


def make_ctx_from_global(lib, mesh_nx, block_nx, xmin, xmax, ng, mhd, gamma, U, B=(None, None, None),
                         flux=None, S=None, bcs=None, grav_acc=None, arith=0):
    mesh = pyhost.UniformMesh(mesh_nx, block_nx, xmin, xmax, ng=ng, bcs=bcs)
    ctx = pyhost.Context(mesh, mhd, gamma, rsolver=RSOLVERS[flux] if flux else None, lib=lib,
                         nscalars=0 if S is None else S.shape[0], grav_acc=grav_acc, arith=arith)
    if S is not None:
        for lb, blk in enumerate(split_to_blocks(mesh, S)):
            ctx.upload("S", lb, blk["u"])
    blocks = split_to_blocks(mesh, U, *B)
    for lb, blk in enumerate(blocks):
        ctx.upload("U", lb, blk["u"])
        if mhd:
            ctx.upload("B_X1F", lb, blk["b1"])
            ctx.upload("B_X2F", lb, blk["b2"])
            ctx.upload("B_X3F", lb, blk["b3"])
    return ctx


def run_golden_case(lib, name, block=None, fused=False, arith=0):
    """Run the driver from a golden fixture's initial state to its tlim. Returns
    (meta, z, ctx, integrator)."""
    meta, z = load_golden(name)
    mhd = meta["mhd"]
    B = (z["b1_0"], z["b2_0"], z["b3_0"]) if mhd else (None, None, None)
    ctx = make_ctx_from_global(lib, meta["mesh"], block or meta["block"], meta["xmin"], meta["xmax"],
                               meta["ng"], mhd, meta["gamma"], z["u0"], B, flux=meta.get("flux"),
                               S=z["s0"] if meta.get("nscalars") else None, bcs=meta.get("bcs"),
                               grav_acc=meta.get("grav_acc"), arith=arith)
    ti = pyhost.TimeIntegrator(ctx, meta["integrator"], meta["xorder"], meta["cfl"], fused=fused)
    ti.initialize(tlim=meta["tlim"])
    ti.run(meta["tlim"])
    return meta, z, ctx, ti


def assert_matches_golden(meta, z, ctx, ti):
    U, B = gather_from_ctx(ctx)
    hst = z["hst"]
    assert ti.ncycle == meta["ncycle"], (ti.ncycle, meta["ncycle"])
    # dt sequence: .hst column 2 holds the dt of every cycle (data_format %24.16e round-trips)
    np.testing.assert_array_equal(np.array(ti.dt_history), hst[:-1, 1][:len(ti.dt_history)])
    assert ti.dt == meta["dt_final"]
    # integer/bit-exact bar: the reference arithmetic is reproduced operation for operation
    np.testing.assert_array_equal(U, z["uf"])
    if meta["mhd"]:
        np.testing.assert_array_equal(B[0], z["b1_f"])
        np.testing.assert_array_equal(B[1], z["b2_f"])
        np.testing.assert_array_equal(B[2], z["b3_f"])
    if meta.get("nscalars"):
        np.testing.assert_array_equal(gather_from_ctx(ctx, "S")[0], z["sf"])
    # conserved-quantity history (north star: 1e-10 relative): athk_history against the last row
    # of the reference's .hst (columns: time, dt, then the sums of outputs/history.cpp:196-248)
    import ctypes
    nv = 8 + (3 if meta["mhd"] else 0) + (meta.get("nscalars") or 0)
    sums = (ctypes.c_double * nv)()
    ctx.call("athk_history", sums, nv)
    ref = hst[-1, 2:2 + nv]
    scale = np.abs(hst[:, 2:2 + nv]).max(axis=0) + 1e-300
    # momenta cancel to ~0 in these periodic problems: scale them by the sum of |terms| instead
    vol = np.prod((np.array(meta["xmax"]) - np.array(meta["xmin"])) / np.array(meta["mesh"]))
    scale[1:4] = vol * np.abs(z["uf"][1:4]).sum(axis=(1, 2, 3)) + 1e-300
    assert np.all(np.abs(np.array(sums[:]) - ref) <= 1e-10 * scale), (np.array(sums[:]), ref)
