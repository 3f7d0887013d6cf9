"""The C++ host (Mesh / MeshBlock / ParameterInput / TimeIntegratorTaskList / pgen mirror) run the
way the reference's regression tests run the reference: same athinput -> run the binary -> parse
its .rst / .hst.  Restart dumps (initial = pgen output, intermediate, final), the dt sequence and
the history file must be IDENTICAL to the reference binary's (oracle/_ref, built from
/root/reference by oracle/Makefile.ref).  CPU variant: host linked with the tests/emul harness
library; GPU variant: the product binary athena_b200/bin/athena_b200 (libathk.so, CUDA)."""
import glob
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_BIN = os.path.join(ROOT, "tests", "emul", "athena_emul")
CUDA_BIN = os.path.join(ROOT, "athena_b200", "bin", "athena_b200")


def blast_blocks(n=32, blk=16, tlim=0.02):
    return {
        "job": {"problem_id": "Blast"},
        "output1": {"file_type": "hst", "dt": 0.005, "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 0.01},
        "time": {"cfl_number": 0.3, "nlim": -1, "tlim": tlim, "integrator": "vl2", "xorder": 2,
                 "ncycle_out": 1},
        "mesh": {"nx1": n, "x1min": -1.0, "x1max": 1.0, "ix1_bc": "periodic", "ox1_bc": "periodic",
                 "nx2": n, "x2min": -1.0, "x2max": 1.0, "ix2_bc": "periodic", "ox2_bc": "periodic",
                 "nx3": n, "x3min": -1.0, "x3max": 1.0, "ix3_bc": "periodic", "ox3_bc": "periodic",
                 "num_threads": 1},
        "meshblock": {"nx1": blk, "nx2": blk, "nx3": blk},
        "hydro": {"gamma": 1.666666666667, "iso_sound_speed": 0.4082482905},
        "problem": {"pamb": 0.1, "prat": 100.0, "radius": 0.1, "damb": 1.0, "drat": 1.0,
                    "ramp": 0.0, "b0": 1.0, "angle": 30},
    }


def sr_blast_blocks(integrator="vl2", xorder=2):
    b = blast_blocks(tlim=0.2)
    b["output1"]["dt"] = 0.05
    b["output2"]["dt"] = 0.1
    b["time"].update(integrator=integrator, xorder=xorder)
    b["problem"].update(prat=1.0e4, pamb=0.01)
    del b["hydro"]["iso_sound_speed"]
    return b


SR_TUBES = {   # inputs/hydro_sr/athinput.mb_N: (gamma, dl, pl, ul, vl, wl, dr, pr, ur, vr, wr)
    1: (1.3333333333333333, 1.0, 1.0, 0.9, 0.0, 0.0, 1.0, 10.0, 0.0, 0.0, 0.0),
    2: (1.6666666666666667, 1.0, 10.0, -0.6, 0.0, 0.0, 10.0, 20.0, 0.5, 0.0, 0.0),
    4: (1.6666666666666667, 1.0, 1.0e3, 0.0, 0.0, 0.0, 1.0, 1.0e-2, 0.0, 0.0, 0.0),
}


def sr_walls_blocks():
    b = refrun.blast_blocks(mesh=(32, 32, 32), block=(16, 16, 16), tlim=0.6, rst_dt=0.3, b0=None, radius=0.3,
                            integrator="vl2", bcs=["reflecting"] * 4 + ["outflow"] * 2, center=(0.5, 0.4, 0.3))
    b["problem"].update(prat=1.0e3, pamb=0.01)
    b["hydro"].pop("iso_sound_speed", None)
    return b


def sr_tube_blocks(which, shock_dir, integrator="vl2", sfloor=None):
    gam, dl, pl, ul, vl, wl, dr, pr, ur, vr, wr = SR_TUBES[which]
    n = [8, 8, 8]
    n[shock_dir - 1] = 64
    blk = [8, 8, 8]
    blk[shock_dir - 1] = 16
    bc = ["periodic"] * 3
    bc[shock_dir - 1] = "outflow"
    return {
        "job": {"problem_id": "srtube"},
        "output1": {"file_type": "hst", "dt": 0.1, "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 0.2},
        "time": {"cfl_number": 0.4, "nlim": -1, "tlim": 0.4, "integrator": integrator, "xorder": 2, "ncycle_out": 1},
        "mesh": {"nx1": n[0], "x1min": -0.5, "x1max": 0.5, "ix1_bc": bc[0], "ox1_bc": bc[0],
                 "nx2": n[1], "x2min": -0.5, "x2max": 0.5, "ix2_bc": bc[1], "ox2_bc": bc[1],
                 "nx3": n[2], "x3min": -0.5, "x3max": 0.5, "ix3_bc": bc[2], "ox3_bc": bc[2], "num_threads": 1},
        "meshblock": {"nx1": blk[0], "nx2": blk[1], "nx3": blk[2]},
        "hydro": {"gamma": gam} if sfloor is None else {"gamma": gam, "sfloor": sfloor},
        "problem": {"shock_dir": shock_dir, "xshock": 0.0, "dl": dl, "pl": pl, "ul": ul, "vl": vl, "wl": wl,
                    "dr": dr, "pr": pr, "ur": ur, "vr": vr, "wr": wr},
    }


def sr_linwave_blocks(wave_flag, integrator="vl2"):
    # wave speeds of hydro_convergence.py's wavespeeds_hydro() for this background; tlim = one period
    gam, rho, pgas, vx, vy, vz = 4.0 / 3.0, 4.0, 1.0, 0.1, 0.3, -0.05
    wgas = rho + gam / (gam - 1.0) * pgas
    cs_sq = gam * pgas / wgas
    v_sq = vx * vx + vy * vy + vz * vz
    u0_sq = 1.0 / (1.0 - v_sq)
    delta = u0_sq * (1.0 - cs_sq) + cs_sq
    a, b = vx * cs_sq, (cs_sq * (u0_sq * (1.0 - cs_sq) * (1.0 - vx * vx) + cs_sq)) ** 0.5
    lam = {0: vx - (a + b) / delta, 4: vx - (a - b) / delta}.get(wave_flag, vx)
    tlim = float(0.5 / abs(lam))   # (half a period, to keep the CPU suite short; the error row is compared as text)
    if tlim > 2.0:
        tlim = 2.0   # (entropy waves drift at vx = 0.1: a fraction of a period is enough for parity)
    return {
        "job": {"problem_id": "srwave"},
        "output1": {"file_type": "hst", "dt": tlim / 2, "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": tlim / 2},
        "time": {"cfl_number": 0.8, "nlim": -1, "tlim": tlim, "integrator": integrator, "xorder": 2, "ncycle_out": 1},
        "mesh": {"nx1": 64, "x1min": -0.5, "x1max": 0.5, "ix1_bc": "periodic", "ox1_bc": "periodic",
                 "nx2": 8, "x2min": -0.5, "x2max": 0.5, "ix2_bc": "periodic", "ox2_bc": "periodic",
                 "nx3": 8, "x3min": -0.5, "x3max": 0.5, "ix3_bc": "periodic", "ox3_bc": "periodic", "num_threads": 1},
        "meshblock": {"nx1": 16, "nx2": 8, "nx3": 8},
        "hydro": {"gamma": gam},
        "problem": {"wave_flag": wave_flag, "amp": 1.0e-6, "compute_error": True, "rho": rho, "pgas": pgas,
                    "vx": vx, "vy": vy, "vz": vz},
    }


def tube_error_blocks(mhd, shock_dir, integrator="vl2"):
    n = [8, 8, 8]
    n[shock_dir - 1] = 64
    blk = [8, 8, 8]
    blk[shock_dir - 1] = 32
    bcs = ["periodic"] * 6
    bcs[2 * (shock_dir - 1)] = bcs[2 * (shock_dir - 1) + 1] = "outflow"
    if mhd:    # Ryu & Jones 2a (inputs/mhd/athinput.rj2a)
        left = (1.08, 1.2, 0.01, 0.5, 0.95, 1.0155412503859613, 0.5641895835477562)
        right = (1.0, 0.0, 0.0, 0.0, 1.0, 1.1283791670955125, 0.5641895835477562)
        b = refrun.shocktube_blocks(True, left, right, shock_dir=shock_dir, bx=0.5641895835477562, xshock=0.0,
                                    mesh=tuple(n), block=tuple(blk), tlim=0.2, rst_dt=0.1, integrator=integrator,
                                    bcs=bcs)
        b["hydro"]["gamma"] = 1.666666666667
    else:      # Sod (inputs/hydro/athinput.sod)
        b = refrun.shocktube_blocks(False, (1.0, 0.0, 0.0, 0.0, 1.0), (0.125, 0.0, 0.0, 0.0, 0.1),
                                    shock_dir=shock_dir, xshock=0.0, mesh=tuple(n), block=tuple(blk), tlim=0.25,
                                    rst_dt=0.125, integrator=integrator, bcs=bcs)
        b["hydro"]["gamma"] = 1.4
    for d in "123":
        b["mesh"]["x%smin" % d], b["mesh"]["x%smax" % d] = -0.5, 0.5
    b["problem"]["compute_error"] = "true"
    return b


def kh_random_blocks(b0):
    b = refrun.kh_blocks(mesh=(16, 32, 8), block=(8, 16, 8), tlim=0.05, rst_dt=0.025, integrator="vl2", xorder=2,
                         iprob=1, vflow=0.5, b0=b0)
    b["problem"]["drat"] = 2.0
    b["mesh"]["x2min"], b["mesh"]["x2max"] = -0.5, 0.5
    b["hydro"]["gamma"] = 1.4
    return b


def with_gravity(blocks, **acc):
    blocks["hydro"].update(acc)
    return blocks


def with_data_dumps(blocks, tab_vars, vtk_vars, ghost_vars):
    dt = blocks["output2"]["dt"]
    blocks["output3"] = {"file_type": "tab", "variable": tab_vars, "dt": dt, "data_format": "%24.16e"}
    blocks["output4"] = {"file_type": "vtk", "variable": vtk_vars, "dt": dt}
    blocks["output5"] = {"file_type": "tab", "variable": ghost_vars, "dt": 2 * dt, "ghost_zones": "true",
                         "id": "withghosts"}
    blocks["output6"] = {"file_type": "vtk", "variable": ghost_vars, "dcycle": 3, "ghost_zones": "true"}
    return blocks


def jet_blocks(mhd, integrator="vl2"):
    """inputs/hydro/athinput.jet in 3-D at test size: user-enrolled inner-x1 boundary (the inlet),
    outflow elsewhere; 2x2x2 MeshBlocks, so the inlet crosses block edges and the boundary function
    sees ghost zones the neighbour exchange filled."""
    b = {
        "job": {"problem_id": "jet"},
        "output1": {"file_type": "hst", "dt": 0.25, "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 0.5},
        "time": {"cfl_number": 0.3, "nlim": -1, "tlim": 1.0, "integrator": integrator, "xorder": 2,
                 "ncycle_out": 1},
        "mesh": {"nx1": 24, "x1min": 0.0, "x1max": 6.0, "ix1_bc": "user", "ox1_bc": "outflow",
                 "nx2": 16, "x2min": -2.0, "x2max": 2.0, "ix2_bc": "outflow", "ox2_bc": "outflow",
                 "nx3": 16, "x3min": -2.0, "x3max": 2.0, "ix3_bc": "outflow", "ox3_bc": "outflow",
                 "num_threads": 1},
        "meshblock": {"nx1": 12, "nx2": 8, "nx3": 8},
        "hydro": {"gamma": 1.666666666667},
        "problem": {"d": 1.0, "p": 6.0e-3, "vx": 0.0, "vy": 0.0, "vz": 0.0, "djet": 0.1, "pjet": 6.0e-3,
                    "vxjet": 1.0, "vyjet": 0.0, "vzjet": 0.0, "rjet": 1.0},
    }
    if mhd:
        b["problem"].update({"bx": 0.1, "by": 0.0, "bz": 0.0, "bxjet": 0.1, "byjet": 0.02, "bzjet": 0.0})
    return b


def refined(blocks, *regions):
    """<mesh> refinement = static with <refinementN> blocks (x1min, x1max, x2min, x2max, x3min, x3max, level)"""
    blocks["mesh"]["refinement"] = "static"
    for n, r in enumerate(regions):
        blocks["refinement%d" % (n + 1)] = dict(zip(("x1min", "x1max", "x2min", "x2max", "x3min", "x3max", "level"), r))
    return blocks


def usrsrc_blocks(mhd, integrator):
    b = {
        "job": {"problem_id": "usrsrc"},
        "output1": {"file_type": "hst", "dt": 0.05, "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 0.1},
        "time": {"cfl_number": 0.3, "nlim": -1, "tlim": 0.2, "integrator": integrator, "xorder": 2, "ncycle_out": 1},
        "mesh": {"nx1": 16, "x1min": -0.5, "x1max": 0.5, "ix1_bc": "periodic", "ox1_bc": "periodic",
                 "nx2": 16, "x2min": -0.5, "x2max": 0.5, "ix2_bc": "periodic", "ox2_bc": "periodic",
                 "nx3": 24, "x3min": -0.75, "x3max": 0.75, "ix3_bc": "periodic", "ox3_bc": "periodic",
                 "num_threads": 1},
        "meshblock": {"nx1": 8, "nx2": 8, "nx3": 12},
        "hydro": {"gamma": 1.4},
        "problem": {"amp": 0.2, "g0": 0.8, "tcool": 0.5},
    }
    if mhd:
        b["problem"]["b0"] = 0.3
    return b


def iso(blocks, cs, **problem):
    blocks["hydro"]["iso_sound_speed"] = cs
    blocks["problem"].update(problem)
    return blocks


CASES = {
    # isothermal equation of state (--eos=isothermal: NHYDRO = 4; HLLE for hydro, hlld_iso for MHD)
    "iso_blast_hydro_hlle": ("blast_iso_hydro_hlle", "blast", False, lambda: iso(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(16, 16, 16), tlim=0.1, rst_dt=0.05, b0=None, radius=0.3),
        0.7, drat=8.0), ["--eos=isothermal"]),
    "iso_blast_hydro_reflecting_rk3": ("blast_iso_hydro_hlle", "blast", False, lambda: iso(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(16, 16, 16), tlim=0.1, rst_dt=0.05, b0=None, radius=0.3,
                            integrator="rk3", bcs=["reflecting"] * 4 + ["outflow"] * 2, center=(0.5, 0.4, 0.3)),
        0.7, drat=8.0), ["--eos=isothermal"]),
    "iso_kh_mhd_hlld_scalar": ("kh_iso_mhd_hlld_s1", "kh", True, lambda: iso(
        refrun.kh_blocks(mesh=(32, 32, 8), block=(16, 16, 8), tlim=0.1, rst_dt=0.05, integrator="vl2", xorder=2,
                         b0=0.5), 1.1), ["--eos=isothermal", "--nscalars=1"]),
    # ... with the reference's other solvers for it: the NON_BAROTROPIC_EOS == 0 branches of llf.cpp, roe.cpp
    # (four waves, LLF fallback), llf_mhd.cpp and hlle_mhd.cpp
    "iso_blast_hydro_llf": ("blast_iso_hydro_llf", "blast", False, lambda: iso(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(16, 16, 16), tlim=0.1, rst_dt=0.05, b0=None, radius=0.3),
        0.7, drat=8.0), ["--eos=isothermal", "--flux=llf"]),
    "iso_blast_hydro_roe_rk3": ("blast_iso_hydro_roe", "blast", False, lambda: iso(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(16, 16, 16), tlim=0.1, rst_dt=0.05, b0=None, radius=0.3,
                            integrator="rk3", bcs=["reflecting"] * 4 + ["outflow"] * 2, center=(0.5, 0.4, 0.3)),
        0.7, drat=30.0), ["--eos=isothermal", "--flux=roe"]),
    "iso_orszag_tang_hlle": ("ot_iso_mhd_hlle", "orszag_tang", True, lambda: iso(
        refrun.ot_blocks(tlim=0.05, rst_dt=0.025), 0.9), ["--eos=isothermal", "--flux=hlle"]),
    "iso_orszag_tang_llf_rk2": ("ot_iso_mhd_llf", "orszag_tang", True, lambda: iso(
        refrun.ot_blocks(tlim=0.05, rst_dt=0.025, integrator="rk2"), 0.9), ["--eos=isothermal", "--flux=llf"]),
    # SURVEY 8(f).4, SR part: special-relativistic hydro (configure.py -s; adiabatic_hydro_sr.cpp's
    # fixed-point W(U) started from W of two stages ago, hlle_rel.cpp, dt = cfl * min(dx)); a strong
    # blast (pressure ratio 1e4), VL2 + PLM and RK3 + PLM
    "sr_blast_hlle": ("blast_sr_hydro_hlle", "blast", False, lambda: sr_blast_blocks(), ["-s", "--flux=hlle"]),
    "sr_blast_hlle_rk3": ("blast_sr_hydro_hlle", "blast", False, lambda: sr_blast_blocks(
        integrator="rk3"), ["-s", "--flux=hlle"]),
    "sr_blast_hllc": ("blast_sr_hydro_hllc", "blast", False, lambda: sr_blast_blocks(), ["-s"]),   # -s default: hllc
    "sr_blast_hllc_rk2": ("blast_sr_hydro_hllc", "blast", False, lambda: sr_blast_blocks(integrator="rk2"),
                          ["-s", "--flux=hllc"]),
    # the reference's own SR regression problems (tst/regression/scripts/tests/sr/hydro_shocks_*.py:
    # inputs/hydro_sr/athinput.mb_1, mb_2, mb_4 -- Mignone & Bodo 2005 tubes, inflow at 0.9c, pressure ratio 1e5), as 3-D slabs with outflow walls along the tube: the pgen seeds w and w1, the walls
    # go through the SR U(W) on the ghost slabs
    "sr_tube_mb1_x1": ("grtube_sr_hydro_hllc", "gr_shock_tube", False, lambda: sr_tube_blocks(1, 1), ["-s"]),
    "sr_tube_mb2_x2": ("grtube_sr_hydro_hllc", "gr_shock_tube", False, lambda: sr_tube_blocks(2, 2), ["-s"]),
    "sr_tube_mb4_x3_rk3": ("grtube_sr_hydro_hllc", "gr_shock_tube", False, lambda: sr_tube_blocks(4, 3, "rk3"), ["-s"]),
    "sr_tube_mb2_x1_scalar_hlle_rk2": ("grtube_sr_hydro_hlle_s1", "gr_shock_tube", False, lambda: sr_tube_blocks(2, 1, "rk2", sfloor=0.0),
                                       ["-s", "--flux=hlle", "--nscalars=1"]),
    # tst/regression/scripts/tests/sr/hydro_convergence.py: linear waves on a moving background
    # (rho 4, pgas 1, v = (0.1, 0.3, -0.05), gamma 4/3, amp 1e-6), half a period, error row appended by
    # UserWorkAfterLoop; a sound wave and an entropy (shear) wave
    "sr_linwave_sound": ("grlinwave_sr_hydro_hllc", "gr_linear_wave", False, lambda: sr_linwave_blocks(0), ["-s"]),
    "sr_linwave_shear_rk3": ("grlinwave_sr_hydro_hllc", "gr_linear_wave", False, lambda: sr_linwave_blocks(2, "rk3"), ["-s"]),
    "sr_blast_hllc_rk3_ppm": ("blast_sr_hydro_hllc_g3", "blast", False, lambda: sr_blast_blocks(integrator="rk3", xorder=3),
                              ["-s", "--nghost=3"]),
    # SR blast off-centre in a box with four reflecting and two outflow walls (primitives mirrored / copied,
    # then the SR U(W) on the ghost slabs)
    "sr_blast_hllc_reflecting": ("blast_sr_hydro_hllc", "blast", False, lambda: sr_walls_blocks(), ["-s"]),
    # .tab / .vtk dumps of an SR run (primitives = four-velocity components, taken from the array W(U) wrote last)
    "sr_tube_mb2_x1_tab_vtk": ("grtube_sr_hydro_hllc", "gr_shock_tube", False, lambda: with_data_dumps(
        sr_tube_blocks(2, 1), "prim", "cons", "prim"), ["-s"]),
    "sr_blast_llf": ("blast_sr_hydro_llf", "blast", False, lambda: sr_blast_blocks(), ["-s", "--flux=llf"]),
    "iso_orszag_tang_roe": ("ot_iso_mhd_roe", "orszag_tang", True, lambda: iso(
        refrun.ot_blocks(tlim=0.08, rst_dt=0.04), 0.6), ["--eos=isothermal", "--flux=roe"]),
    "iso_orszag_tang_rk3": ("ot_iso_mhd_hlld", "orszag_tang", True, lambda: iso(
        refrun.ot_blocks(tlim=0.05, rst_dt=0.025, integrator="rk3"), 0.9), ["--eos=isothermal"]),
    # user-enrolled explicit source terms (SrcTermFunc, time-dependent) -- the reference
    # side is the unmodified reference linked with the user problem file oracle/pgen/usrsrc.cpp
    "user_source_hydro_rk3": ("usrsrc_hydro_hllc", "usrsrc", False, lambda: usrsrc_blocks(False, "rk3")),
    "user_source_mhd_vl2": ("usrsrc_mhd_hlld", "usrsrc", True, lambda: usrsrc_blocks(True, "vl2")),
    # static mesh refinement, hydro (restriction / prolongation of the cell-centred
    # variables, flux correction).  The reference's own linear-wave regression runs refined.
    "smr_linwave_hydro_centre": ("linwave_hydro_hllc", "linear_wave", False, lambda: refined(
        refrun.linwave_blocks(False, mesh=(32, 16, 16), block=(8, 8, 8), tlim=0.05, rst_dt=0.025),
        (0.75, 2.25, 0.375, 1.125, 0.375, 1.125, 1))),
    "smr_linwave_hydro_two_levels_rk3": ("linwave_hydro_hllc", "linear_wave", False, lambda: refined(
        refrun.linwave_blocks(False, mesh=(24, 16, 16), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, integrator="rk3"),
        (0.0, 0.6, 0.0, 0.5, 0.9, 1.5, 2))),   # region in a corner of the periodic box, 2:1 balance cascades
    "smr_blast_hydro": ("blast_hydro_hllc", "blast", False, lambda: refined(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(8, 8, 8), tlim=0.06, rst_dt=0.03, b0=None, radius=0.25),
        (-0.3, 0.3, -0.3, 0.3, -0.3, 0.3, 1))),
    # ... with walls: the boundary functions also act on the coarse buffer (bvals_refine.cpp:445-530);
    # one region in a reflecting corner, a two-level one in an outflow corner
    "smr_blast_hydro_walls": ("blast_hydro_hllc", "blast", False, lambda: refined(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(8, 8, 8), tlim=0.08, rst_dt=0.04, b0=None, radius=0.25,
                            bcs=["reflecting", "outflow", "outflow", "reflecting", "reflecting", "outflow"],
                            center=(-0.6, 0.5, -0.55)),
        (-1.0, -0.4, 0.3, 1.0, -1.0, -0.3, 1), (0.5, 1.0, -1.0, -0.6, 0.4, 1.0, 2))),
    "smr_linwave_hydro_ppm_two_regions": ("linwave_hydro_hllc_g4", "linear_wave", False, lambda: refined(
        refrun.linwave_blocks(False, mesh=(32, 16, 16), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, integrator="rk3",
                              xorder=3),
        (1.0, 1.6, 0.2, 0.5, 0.9, 1.2, 1), (2.4, 2.9, 1.0, 1.4, 0.1, 0.3, 1)), ["--nghost=4"]),
    # ... and MHD: restriction / prolongation of the face-centred field (shared faces, then the
    # divergence-preserving internal faces), same-level / to-coarser / to-finer field buffers with the
    # faces edge and corner neighbours share, EMF correction from finer blocks.  One level in the
    # middle of the box; two levels in a corner of the periodic box with RK3; a blast crossing the
    # level boundary; PPM with nghost=4 (cnghost=3) and two separate regions
    "smr_linwave_mhd_centre": ("linwave_mhd_hlld", "linear_wave", True, lambda: refined(
        refrun.linwave_blocks(True, mesh=(32, 16, 16), block=(8, 8, 8), tlim=0.05, rst_dt=0.025),
        (0.75, 2.25, 0.375, 1.125, 0.375, 1.125, 1))),
    "smr_linwave_mhd_two_levels_rk3": ("linwave_mhd_hlld", "linear_wave", True, lambda: refined(
        refrun.linwave_blocks(True, mesh=(24, 16, 16), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, integrator="rk3"),
        (0.0, 0.6, 0.0, 0.5, 0.9, 1.5, 2))),
    "smr_blast_mhd": ("blast_mhd_hlld", "blast", True, lambda: refined(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, radius=0.25),
        (-0.3, 0.3, -0.3, 0.3, -0.3, 0.3, 1))),
    "smr_linwave_mhd_ppm_two_regions": ("linwave_mhd_hlld_g4", "linear_wave", True, lambda: refined(
        refrun.linwave_blocks(True, mesh=(32, 16, 16), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, integrator="rk3",
                              xorder=3),
        (1.0, 1.6, 0.2, 0.5, 0.9, 1.2, 1), (2.4, 2.9, 1.0, 1.4, 0.1, 0.3, 1)), ["--nghost=4"]),
    # ... between walls: the reflecting / outflow functions of the field on the coarse buffer
    "smr_blast_mhd_walls": ("blast_mhd_hlld", "blast", True, lambda: refined(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, radius=0.25,
                            bcs=["reflecting", "outflow", "outflow", "reflecting", "reflecting", "outflow"],
                            center=(-0.6, 0.5, -0.55)),
        (-1.0, -0.4, 0.3, 1.0, -1.0, -0.3, 1), (0.5, 1.0, -1.0, -0.6, 0.4, 1.0, 2))),
    # ... with a user-enrolled boundary function (the jet's inlet) acting on the coarse buffer too
    "smr_jet_user_bc_hydro": ("jet_hydro_hllc", "jet", False, lambda: refined(
        jet_blocks(False), (0.0, 1.0, 0.0, 2.0, -1.2, 0.5, 1))),
    "smr_jet_user_bc_mhd_rk3": ("jet_mhd_hlld", "jet", True, lambda: refined(
        jet_blocks(True, "rk3"), (0.0, 1.0, -2.0, 0.0, -0.5, 1.2, 1))),
    # ... with passive scalars (restriction of s, prolongation of r = s / rho, flux correction of s_flux)
    "smr_kh_hydro_2scalars": ("kh_hydro_hllc_s2", "kh", False, lambda: refined(
        refrun.kh_blocks(mesh=(16, 32, 16), block=(8, 8, 8), tlim=0.05, rst_dt=0.025, integrator="vl2", xorder=2,
                         drho_rho0=0.5), (-0.2, 0.2, 0.3, 0.7, -0.1, 0.1, 1)), ["--nscalars=2"]),
    "smr_kh_mhd_scalar_rk3": ("kh_mhd_hlld_s1", "kh", True, lambda: refined(
        refrun.kh_blocks(mesh=(16, 32, 16), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, integrator="rk3", xorder=2,
                         b0=0.5), (-0.5, 0.0, -0.7, -0.3, -0.25, 0.0, 1)), ["--nscalars=1"]),
    # ... with the isothermal EOS (NHYDRO = 4): hydro blast with walls, Orszag-Tang vortex (hlld_iso) refined
    # off-centre (its potential is sampled per level, as in the reference: the run is the comparison)
    "smr_iso_blast_hydro_walls": ("blast_iso_hydro_hlle", "blast", False, lambda: refined(iso(
        refrun.blast_blocks(mesh=(32, 32, 32), block=(8, 8, 8), tlim=0.08, rst_dt=0.04, b0=None, radius=0.3,
                            bcs=["reflecting"] * 2 + ["outflow"] * 2 + ["periodic"] * 2, center=(-0.5, 0.4, 0.3)),
        0.7, drat=8.0), (-1.0, -0.3, 0.2, 1.0, 0.0, 0.6, 1)), ["--eos=isothermal"]),
    "smr_iso_orszag_tang": ("ot_iso_mhd_hlld", "orszag_tang", True, lambda: refined(iso(
        refrun.ot_blocks(mesh=(32, 32, 16), block=(8, 8, 8), tlim=0.04, rst_dt=0.02), 0.9),
        (-0.2, 0.3, -0.3, 0.1, 0.0, 0.2, 1)), ["--eos=isothermal"]),
    # user-enrolled boundary functions (Mesh::EnrollUserBoundaryFunction, BValFunc)
    "jet_user_bc_hydro": ("jet_hydro_hllc", "jet", False, lambda: jet_blocks(False)),
    "jet_user_bc_mhd_rk3": ("jet_mhd_hlld", "jet", True, lambda: jet_blocks(True, "rk3")),
    "mhd_linwave": ("linwave_mhd_hlld", "linear_wave", True,
                    lambda: dict(refrun.linwave_blocks(True, mesh=(32, 16, 16), block=(16, 8, 8),
                                                       tlim=0.05, rst_dt=0.02))),
    "hydro_linwave_rk3": ("linwave_hydro_hllc", "linear_wave", False,
                          lambda: dict(refrun.linwave_blocks(False, mesh=(32, 16, 16), block=(16, 8, 8),
                                                             tlim=0.1, rst_dt=0.05, integrator="rk3"))),
    "mhd_blast": ("blast_mhd_hlld", "blast", True, blast_blocks),
    # the reference's other integrators: RK4()4[2S] (four stages, two registers: fused stage) and
    # SSPRK(5,4) (five stages, third register u2 / s2 and the extra update of its fourth stage)
    "mhd_linwave_rk4": ("linwave_mhd_hlld", "linear_wave", True,
                        lambda: dict(refrun.linwave_blocks(True, mesh=(32, 16, 16), block=(16, 8, 8),
                                                           tlim=0.05, rst_dt=0.02, integrator="rk4"))),
    "kh_ssprk5_4_2scalars": ("kh_hydro_hllc_s2", "kh", False, lambda: refrun.kh_blocks(
        mesh=(16, 16, 8), block=(8, 8, 8), tlim=0.04, rst_dt=0.015, integrator="ssprk5_4", xorder=2), ["--nscalars=2"]),
    # BASELINE configs[3] at parity size: KH + RK3 + PPM + HLLC + one passive scalar (row a10)
    "kh_rk3_ppm_scalar": ("kh_hydro_hllc_s1g3", "kh", False,
                          lambda: refrun.kh_blocks(mesh=(32, 32, 8), block=(16, 16, 8), tlim=0.03,
                                                   rst_dt=0.015, drho_rho0=0.5),
                          ["--nghost=3", "--nscalars=1"]),
    "kh_vl2_plm_2scalars": ("kh_hydro_hllc_s2", "kh", False,
                            lambda: refrun.kh_blocks(mesh=(32, 32, 8), block=(16, 16, 8), tlim=0.03,
                                                     rst_dt=0.015, integrator="vl2", xorder=2),
                            ["--nscalars=2"]),
    "kh_mhd_rk2_scalar": ("kh_mhd_hlld_s1", "kh", True,
                          lambda: refrun.kh_blocks(mesh=(32, 32, 8), block=(16, 16, 8), tlim=0.03,
                                                   rst_dt=0.015, integrator="rk2", xorder=2, b0=0.5),
                          ["--nscalars=1"]),
    # physical boundaries through the host's <mesh> ix*_bc / ox*_bc flags
    "sod_outflow_x1": ("shocktube_hydro_hllc", "shock_tube", False,
                       lambda: refrun.shocktube_blocks(False, (1.0, 0.0, 0.0, 0.0, 1.0), (0.125, 0.0, 0.0, 0.0, 0.1),
                                                       shock_dir=1, mesh=(32, 16, 16), block=(16, 8, 8), tlim=0.1,
                                                       rst_dt=0.05, bcs=["outflow"] * 2 + ["periodic"] * 4)),
    "briowu_x3_reflect_outflow": ("shocktube_mhd_hlld", "shock_tube", True,
                                  lambda: refrun.shocktube_blocks(
                                      True, (1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0), (0.125, 0.0, 0.0, 0.0, 0.1, -1.0, 0.0),
                                      shock_dir=3, bx=0.75, mesh=(32, 16, 16), block=(16, 8, 8), tlim=0.05,
                                      rst_dt=0.025, bcs=["periodic"] * 4 + ["reflecting", "outflow"])),
    "mhd_blast_reflecting_box": ("blast_mhd_hlld", "blast", True,
                                 lambda: refrun.blast_blocks(tlim=0.06, rst_dt=0.03, bcs=["reflecting"] * 6,
                                                             center=(0.6, 0.55, 0.5), radius=0.3)),
    # BASELINE configs[2] names Orszag-Tang next to the blast wave
    "orszag_tang_rk2": ("ot_mhd_hlld", "orszag_tang", True,
                        lambda: refrun.ot_blocks(tlim=0.05, rst_dt=0.025, integrator="rk2")),
    # the two classic CT problems of the reference's MHD regression set (mhd/cpaw.py; Gardiner & Stone
    # field loop): vector-potential initial data, errors appended by UserWorkAfterLoop
    "cpaw_standing_rk3": ("cpaw_mhd_hlld", "cpaw", True,
                          lambda: refrun.cpaw_blocks(tlim=0.2, rst_dt=0.1, integrator="rk3", v_par=1.0)),
    "cpaw_travelling_left": ("cpaw_mhd_hlld", "cpaw", True,
                             lambda: refrun.cpaw_blocks(tlim=0.2, rst_dt=0.1, direction=2)),
    "field_loop_x1x2": ("fieldloop_mhd_hlld", "field_loop", True,
                        lambda: refrun.fieldloop_blocks(iprob=1, tlim=0.1, rst_dt=0.05)),
    "field_loop_rotated": ("fieldloop_mhd_hlld", "field_loop", True,
                           lambda: refrun.fieldloop_blocks(iprob=4, tlim=0.1, rst_dt=0.05, integrator="rk2")),
    "field_loop_sphere_dense": ("fieldloop_mhd_hlld", "field_loop", True,
                                lambda: refrun.fieldloop_blocks(iprob=5, tlim=0.1, rst_dt=0.05, drat=2.0)),
    # hydro/sod_shock.py and mhd/rj2a_shock.py: tubes on [-0.5, 0.5] with outflow ends, errors against
    # the tabulated exact solutions appended to shock-errors.dat (shock_tube.cpp:46-270)
    "sod_errors_x2": ("shocktube_hydro_hllc", "shock_tube", False, lambda: tube_error_blocks(False, 2)),
    "rj2a_errors_x1": ("shocktube_mhd_hlld", "shock_tube", True, lambda: tube_error_blocks(True, 1)),
    "rj2a_errors_x3_rk3": ("shocktube_mhd_hlld", "shock_tube", True,
                           lambda: tube_error_blocks(True, 3, integrator="rk3")),
    # formatted-table and VTK dumps (outputs/formatted_table.cpp, vtk.cpp): byte-identical files
    "mhd_linwave_tab_vtk": ("linwave_mhd_hlld", "linear_wave", True, lambda: with_data_dumps(
        refrun.linwave_blocks(True, mesh=(16, 8, 8), block=(8, 4, 4), tlim=0.06, rst_dt=0.03), "prim", "cons", "b,d")),
    "kh_scalars_tab_vtk": ("kh_hydro_hllc_s2", "kh", False, lambda: with_data_dumps(
        refrun.kh_blocks(mesh=(16, 16, 8), block=(8, 8, 8), tlim=0.02, rst_dt=0.01, integrator="vl2", xorder=2),
        "cons", "prim", "r1,m2,vx,E"), ["--nscalars=2"]),
    # constant accelerations <hydro> grav_acc1/2/3 (task SRC_TERM, hydro/srcterms/constant_acc.cpp)
    "blast_hydro_gravity_x3_reflect": ("blast_hydro_hllc", "blast", False, lambda: with_gravity(
        refrun.blast_blocks(tlim=0.06, rst_dt=0.03, integrator="rk3",
                            bcs=["periodic"] * 4 + ["reflecting"] * 2, center=(0.1, 0.0, -0.2), radius=0.3),
        grav_acc3=-0.7)),
    "kh_mhd_gravity_all_axes": ("kh_mhd_hlld_s1", "kh", True, lambda: with_gravity(
        refrun.kh_blocks(mesh=(32, 32, 8), block=(16, 16, 8), tlim=0.03, rst_dt=0.015, integrator="vl2", xorder=2,
                         b0=0.5), grav_acc1=0.3, grav_acc2=-0.5, grav_acc3=0.2), ["--nscalars=1"]),
    # kh iprob=1: slip surfaces with white noise from the reference's generator (utils/ran2.cpp), one
    # stream per MeshBlock seeded by its gid; scalars mark stream / ambient medium
    "kh_random_2scalars": ("kh_hydro_hllc_s2", "kh", False, lambda: kh_random_blocks(None), ["--nscalars=2"]),
    "kh_random_mhd": ("kh_mhd_hlld_s1", "kh", True, lambda: kh_random_blocks(0.2), ["--nscalars=1"]),
    # --flux=roe through the host's configure switch
    "mhd_blast_roe": ("blast_mhd_roe", "blast", True, blast_blocks, ["--flux=roe"]),
}


def run_host(binary, blocks, prob, mhd, extra=(), wd=None):
    wd = wd or tempfile.mkdtemp(prefix="athhost_")
    inp = os.path.join(wd, "athinput.host")
    with open(inp, "w") as f:
        f.write(refrun.format_athinput(blocks))
    args = [binary, "-i", inp, "-d", wd, "--prob=" + prob] + (["-b"] if mhd else []) + list(extra)
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    return wd, r


def compare_with_reference(binary, case, host_time_params=None, host_mesh_params=None):
    cfg, prob, mhd, mk = CASES[case][:4]
    extra = CASES[case][4] if len(CASES[case]) > 4 else []
    if not refrun.have_ref(cfg):
        pytest.skip("reference binary oracle/_ref/athena_%s not built" % cfg)
    blocks = mk()
    wd_ref, _ = refrun.run_reference(cfg, blocks)
    host_blocks = {k: dict(v) for k, v in blocks.items()}
    host_blocks["time"].update(host_time_params or {})
    host_blocks["mesh"].update(host_mesh_params or {})
    wd, r = run_host(binary, host_blocks, prob, mhd, extra=extra)
    try:
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        ref_files = sorted(os.path.basename(f) for f in glob.glob(wd_ref + "/*.rst"))
        our_files = sorted(os.path.basename(f) for f in glob.glob(wd + "/*.rst"))
        assert ref_files == our_files and len(ref_files) >= 3
        for name in ref_files:
            a, c = refrun.read_rst(os.path.join(wd_ref, name)), refrun.read_rst(os.path.join(wd, name))
            assert (a["ncycle"], a["time"], a["dt"], a["loc"]) == (c["ncycle"], c["time"], c["dt"], c["loc"]), name
            for k in ("u", "b1", "b2", "b3", "s"):
                assert len(a[k]) == len(c[k])
                for x, y in zip(a[k], c[k]):
                    np.testing.assert_array_equal(x, y, err_msg="%s %s" % (name, k))
        n1, h1 = refrun.read_hst(glob.glob(wd_ref + "/*.hst")[0])
        n2, h2 = refrun.read_hst(glob.glob(wd + "/*.hst")[0])
        assert n1 == n2
        # history sums are formed on the host in the reference's serial order: identical digits;
        # the north-star bar is 1e-10 relative to the column scale
        scale = np.abs(h1).max(axis=0) + 1e-300
        assert np.all(np.abs(h1 - h2) <= 1e-10 * scale)
        np.testing.assert_array_equal(h1[:, :2], h2[:, :2])   # time and dt columns exactly
        for f in glob.glob(wd_ref + "/*-errors.dat"):         # rows UserWorkAfterLoop appends
            assert open(f).read() == open(os.path.join(wd, os.path.basename(f))).read()
        for f in glob.glob(wd_ref + "/*.tab") + glob.glob(wd_ref + "/*.vtk"):   # per-MeshBlock data dumps
            ours = os.path.join(wd, os.path.basename(f))
            assert os.path.exists(ours), os.path.basename(f)
            assert open(f, "rb").read() == open(ours, "rb").read(), os.path.basename(f)
    finally:
        shutil.rmtree(wd, ignore_errors=True)
        shutil.rmtree(wd_ref, ignore_errors=True)


@pytest.fixture(scope="module")
def emul_binary(emul_lib):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "athena_b200", "csrc", "host"), "emul"],
                   check=True, capture_output=True)
    return EMUL_BIN


@pytest.mark.parametrize("case", sorted(CASES))
def test_host_with_emulated_kernels_reproduces_reference_files(emul_binary, case):
    compare_with_reference(emul_binary, case)


@pytest.mark.parametrize("case", ["kh_rk3_ppm_scalar", "kh_vl2_plm_2scalars", "mhd_blast_roe",
                                  "blast_hydro_gravity_x3_reflect", "kh_mhd_gravity_all_axes", "jet_user_bc_mhd_rk3",
                                  "user_source_mhd_vl2", "iso_kh_mhd_hlld_scalar", "iso_blast_hydro_reflecting_rk3",
                                  "sr_blast_hlle"])
def test_host_operator_sequence_reproduces_reference_files(emul_binary, case):
    # time/fused_stage=false: one C-ABI call per reference task instead of athk_stage()
    compare_with_reference(emul_binary, case, host_time_params={"fused_stage": "false"})


def test_ssprk5_4_refuses_magnetic_fields_like_the_reference(emul_binary):
    # time_integrator.cpp:1620-1626: a fatal error at the first stage
    blocks = refrun.linwave_blocks(True, mesh=(16, 8, 8), block=(8, 8, 8), tlim=0.05, integrator="ssprk5_4")
    wd, r = run_host(emul_binary, blocks, "linear_wave", True)
    shutil.rmtree(wd, ignore_errors=True)
    assert r.returncode != 0
    assert "integrator=ssprk5_4 is currently incompatible with MHD." in r.stdout + r.stderr


@pytest.mark.parametrize("case,ncycle", [("smr_linwave_mhd_centre", 1), ("smr_linwave_mhd_two_levels_rk3", 0),
                                         ("smr_blast_mhd_walls", 0), ("smr_linwave_mhd_ppm_two_regions", 0),
                                         ("smr_kh_mhd_scalar_rk3", 0)])
def test_refinement_lists_do_not_depend_on_box_order(emul_binary, case, ncycle):
    """The boxes of one refinement list run concurrently on the device.  The harness library runs every
    list forwards and backwards from the same state (ATHK_REFINE_ORDER_CHECK, tests/emul/
    athk_emul_refine_check.h) and fails if two boxes of a list disagree about an element -- e.g. two
    coarser neighbours handing over their copies of the face they share, which is why the host keeps
    face, edge and corner neighbours in separate lists (refinement.cpp)."""
    cfg, prob, mhd, mk = CASES[case][:4]
    extra = CASES[case][4] if len(CASES[case]) > 4 else []
    blocks = mk()
    blocks["time"]["nlim"] = ncycle   # (0: the exchange of Mesh::Initialize only)
    os.environ["ATHK_REFINE_ORDER_CHECK"] = "1"
    try:
        wd, r = run_host(emul_binary, blocks, prob, mhd, extra=extra)
    finally:
        del os.environ["ATHK_REFINE_ORDER_CHECK"]
    shutil.rmtree(wd, ignore_errors=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "cycle=%d" % ncycle in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_host_binary_on_gpu_reproduces_reference_files(cuda_lib, case):
    assert os.access(CUDA_BIN, os.X_OK), "athena_b200/bin/athena_b200 not built"
    compare_with_reference(CUDA_BIN, case)


# ---- MeshBlocks partitioned over several GPUs by the C++ host (<mesh> num_gpus) ----------------
# Mesh::CalculateLoadBalance ranges + remote links + peer-to-peer transport (host/multigpu.cpp).
# CPU: the harness library treats every "device" as host memory, which exercises the block ranges,
# neighbour tables, message layouts and stage phases; GPU: needs a box with that many devices.
@pytest.mark.parametrize("ngpu,fused", [(2, "true"), (3, "true"), (4, "false")])
def test_host_partitions_blocks_over_gpus_emulated(emul_binary, ngpu, fused):
    def mk():
        b = refrun.linwave_blocks(True, mesh=(32, 16, 16), block=(8, 8, 8), tlim=0.05, rst_dt=0.02)
        return b
    CASES["_mg"] = ("linwave_mhd_hlld", "linear_wave", True, mk)
    try:
        compare_with_reference(emul_binary, "_mg", host_time_params={"fused_stage": fused},
                               host_mesh_params={"num_gpus": ngpu})
    finally:
        del CASES["_mg"]


def test_host_partitions_sr_blocks_over_gpus_emulated(emul_binary):
    # special relativity with the blocks of the tube split over three (emulated) GPUs: ghost messages
    # carry (D, M, E), W(U) on the received ghosts starts from each GPU's own w / w1
    compare_with_reference(emul_binary, "sr_tube_mb2_x2", host_mesh_params={"num_gpus": 3})


@pytest.mark.parametrize("case,ngpu", [("smr_linwave_mhd_two_levels_rk3", 3), ("smr_kh_mhd_scalar_rk3", 4)])
def test_host_partitions_refined_mesh_over_gpus_emulated(emul_binary, case, ngpu):
    """Refined meshes over several GPUs: same-level messages through the remote links, boxes between
    levels (and the field's same-level boxes) read the other context's arrays directly (athk_box_op.peer)."""
    compare_with_reference(emul_binary, case, host_mesh_params={"num_gpus": ngpu})


def _visible_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("case,ngpu", [("mhd_blast", 2), ("kh_mhd_rk2_scalar", 2), ("mhd_blast", 8),
                                       ("smr_linwave_mhd_two_levels_rk3", 2), ("smr_blast_mhd_walls", 2),
                                       ("smr_kh_mhd_scalar_rk3", 8), ("sr_tube_mb2_x2", 2), ("sr_blast_hllc_reflecting", 2)])
def test_host_binary_on_several_gpus_reproduces_reference_files(cuda_lib, case, ngpu):
    if _visible_gpus() < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    assert os.access(CUDA_BIN, os.X_OK), "athena_b200/bin/athena_b200 not built"
    compare_with_reference(CUDA_BIN, case, host_mesh_params={"num_gpus": ngpu})


# ---- parity AT SIZE (GPU only; the reference runs on all host cores) --------------------------
def compare_final_state(binary, cfg, prob, mhd, blocks, threads, errors_file=None):
    """Run reference and host on the same athinput; every .rst (ghost zones included), the dt /
    time columns of the .hst and, where given, the appended error rows must be identical."""
    if not refrun.have_ref(cfg):
        pytest.skip("reference binary oracle/_ref/athena_%s not built" % cfg)
    wd_ref, _ = refrun.run_reference(cfg, blocks, threads=threads)
    wd, r = run_host(binary, blocks, prob, mhd)
    try:
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        ref_files = sorted(os.path.basename(f) for f in glob.glob(wd_ref + "/*.rst"))
        our_files = sorted(os.path.basename(f) for f in glob.glob(wd + "/*.rst"))
        assert ref_files == our_files and len(ref_files) >= 2
        for name in ref_files:
            a, c = refrun.read_rst(os.path.join(wd_ref, name)), refrun.read_rst(os.path.join(wd, name))
            assert (a["ncycle"], a["time"], a["dt"], a["loc"]) == (c["ncycle"], c["time"], c["dt"], c["loc"]), name
            for k in ("u", "b1", "b2", "b3"):
                assert len(a[k]) == len(c[k])
                for x, y in zip(a[k], c[k]):
                    np.testing.assert_array_equal(x, y, err_msg="%s %s" % (name, k))
            ncycle = a["ncycle"]
        _, h1 = refrun.read_hst(glob.glob(wd_ref + "/*.hst")[0])
        _, h2 = refrun.read_hst(glob.glob(wd + "/*.hst")[0])
        np.testing.assert_array_equal(h1[:, :2], h2[:, :2])
        scale = np.abs(h1).max(axis=0) + 1e-300
        assert np.all(np.abs(h1 - h2) <= 1e-10 * scale)
        if errors_file:
            assert open(os.path.join(wd_ref, errors_file)).read() == open(os.path.join(wd, errors_file)).read()
        return ncycle
    finally:
        shutil.rmtree(wd, ignore_errors=True)
        shutil.rmtree(wd_ref, ignore_errors=True)


@pytest.mark.gpu
def test_baseline_config1_at_full_size_on_gpu(cuda_lib):
    """BASELINE configs[1] as stated: MHD linear wave, 128^3 mesh in 32^3 MeshBlocks (64 blocks),
    VL2 + PLM + HLLD + CT, to tlim = 0.5 -- final restart dump bit-equal, linearwave-errors.dat
    text-equal (the L1 error vector, north star 1e-12)."""
    assert os.access(CUDA_BIN, os.X_OK), "athena_b200/bin/athena_b200 not built"
    blocks = refrun.linwave_blocks(True, mesh=(128, 128, 128), block=(32, 32, 32), tlim=0.5, rst_dt=0.5)
    blocks["problem"]["compute_error"] = "true"
    ncycle = compare_final_state(CUDA_BIN, "linwave_mhd_hlld", "linear_wave", True, blocks,
                                 threads=min(os.cpu_count() or 1, 64), errors_file="linearwave-errors.dat")
    assert ncycle > 100


@pytest.mark.gpu
def test_bench_workload_shape_matches_reference_on_gpu(cuda_lib):
    """The shape bench.py times -- MHD blast, 64^3 MeshBlocks, several blocks per context -- against
    the reference: 128^3 mesh (8 blocks of 64^3), 5 VL2 cycles, restart dumps bit-equal."""
    assert os.access(CUDA_BIN, os.X_OK), "athena_b200/bin/athena_b200 not built"
    blocks = refrun.blast_blocks(mesh=(128, 128, 128), block=(64, 64, 64), tlim=1.0, rst_dt=1.0)
    blocks["time"]["nlim"] = 5
    ncycle = compare_final_state(CUDA_BIN, "blast_mhd_hlld", "blast", True, blocks, threads=min(os.cpu_count() or 1, 8))
    assert ncycle == 5


# The reference's convergence regression (tst/regression/scripts/tests/{hydro,mhd}/*_linear_wave*):
# run the wave for one period at two resolutions with problem/compute_error=true and read the rows
# Mesh::UserWorkAfterLoop appends to linearwave-errors.dat (pgen/linear_wave.cpp:190-427).  North
# star: the L1 error vectors agree to 1e-12 relative -- the printed rows are identical text.
LINWAVE_ERR_CASES = {
    "hydro_sound_plm": ("linwave_hydro_hllc", False, dict(wave_flag=0)),
    "hydro_entropy_rk3": ("linwave_hydro_hllc", False, dict(wave_flag=3, vflow=1.0, integrator="rk3")),
    "mhd_fast_plm": ("linwave_mhd_hlld", True, dict(wave_flag=0)),
    "mhd_alfven_rk2": ("linwave_mhd_hlld", True, dict(wave_flag=1, integrator="rk2")),
}


def compare_linear_wave_errors(binary, case):
    cfg, mhd, kw = LINWAVE_ERR_CASES[case]
    if not refrun.have_ref(cfg):
        pytest.skip("reference binary oracle/_ref/athena_%s not built" % cfg)
    wd_ref, wd = tempfile.mkdtemp(prefix="athref_"), tempfile.mkdtemp(prefix="athhost_")
    try:
        for n in (16, 32):
            blocks = refrun.linwave_blocks(mhd, mesh=(n, n // 2, n // 2), block=(n // 2, n // 4, n // 4),
                                           tlim=1.0, **kw)
            blocks["problem"]["compute_error"] = "true"
            refrun.run_reference(cfg, blocks, workdir=wd_ref)
            _, r = run_host(binary, blocks, "linear_wave", mhd, wd=wd)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        ref = open(os.path.join(wd_ref, "linearwave-errors.dat")).read()
        ours = open(os.path.join(wd, "linearwave-errors.dat")).read()
        assert ours == ref        # header once, one row per resolution, every printed digit
        rows = np.array([[float(x) for x in l.split()] for l in ours.splitlines() if not l.startswith("#")])
        assert rows.shape[0] == 2 and rows[0, 0] == 16 and rows[1, 0] == 32
        nv = 8 if mhd else 5
        np.testing.assert_allclose(rows[:, 4], np.sqrt((rows[:, 5:5 + nv] ** 2).sum(axis=1)), rtol=1e-5)
        assert rows[1, 4] < rows[0, 4] / 3.0   # second-order scheme: RMS L1 error drops ~4x per doubling
    finally:
        shutil.rmtree(wd, ignore_errors=True)
        shutil.rmtree(wd_ref, ignore_errors=True)


@pytest.mark.parametrize("case", sorted(LINWAVE_ERR_CASES))
def test_linear_wave_error_vectors_match_reference(emul_binary, case):
    compare_linear_wave_errors(emul_binary, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(LINWAVE_ERR_CASES))
def test_linear_wave_error_vectors_on_gpu_match_reference(cuda_lib, case):
    assert os.access(CUDA_BIN, os.X_OK), "athena_b200/bin/athena_b200 not built"
    compare_linear_wave_errors(CUDA_BIN, case)


# tst/regression/scripts/tests/symmetry/hydro_linwave_aligned.py, restricted to the 3-D update: a
# sound wave along x1, x2 and x3 of a cube.  The x1 sweep is a different kernel from the x2/x3
# sweeps (rows vs shared-memory tiles), so the three runs must agree the way the reference's do:
# identical error rows under the momentum permutation, and (stronger than the reference's 5e-15)
# the same density/energy/normal-momentum profile along the wave vector, digit for digit.
def global_cons(rst):
    nx1, nx2, nx3 = rst["nx"]
    ng = rst["ng"]
    m = rst["mesh_nx"]
    g = np.zeros((5, m[2], m[1], m[0]))
    for (l1, l2, l3, _), u in zip(rst["loc"], rst["u"]):
        g[:, l3 * nx3:(l3 + 1) * nx3, l2 * nx2:(l2 + 1) * nx2, l1 * nx1:(l1 + 1) * nx1] = \
            u[:, ng:ng + nx3, ng:ng + nx2, ng:ng + nx1]
    return g


def check_axis_symmetry(binary, integrator, xorder, extra):
    wd = tempfile.mkdtemp(prefix="athhost_")
    aligned = (dict(ang_2=0.0, ang_3=0.0), dict(ang_2=0.0, ang_3_vert="true"), dict(ang_2_vert="true"))
    try:
        fields = []
        for kw in aligned:
            blocks = refrun.linwave_blocks(False, mesh=(16, 16, 16), block=(8, 8, 8), tlim=1.0, rst_dt=1.0,
                                           integrator=integrator, xorder=xorder)
            for d in "123":
                blocks["mesh"]["x%smax" % d] = 1.0
            blocks["problem"].update(kw, compute_error="true")
            for f in glob.glob(wd + "/*.rst"):
                os.remove(f)
            _, r = run_host(binary, blocks, "linear_wave", False, extra=extra, wd=wd)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            fields.append(global_cons(refrun.read_rst(sorted(glob.glob(wd + "/*.rst"))[-1])))
        rows = [l.split() for l in open(os.path.join(wd, "linearwave-errors.dat")).read().splitlines()[1:]]
        perm = ([6, 7, 8], [7, 8, 6], [8, 6, 7])     # (normal, transverse, transverse) momentum L1 columns
        key = lambda row, p: row[3:6] + [row[c] for c in p] + row[9:12] + [row[c + 6] for c in p] + row[15:]   # noqa: E731
        assert key(rows[0], perm[0]) == key(rows[1], perm[1]) == key(rows[2], perm[2])
        u1, u2, u3 = fields
        for a, b in ((IDN, IDN), (IEN, IEN)):
            np.testing.assert_array_equal(u1[a, 0, 0, :], u2[b, 0, :, 0])
            np.testing.assert_array_equal(u1[a, 0, 0, :], u3[b, :, 0, 0])
        np.testing.assert_array_equal(u1[1, 0, 0, :], u2[2, 0, :, 0])
        np.testing.assert_array_equal(u1[1, 0, 0, :], u3[3, :, 0, 0])
        assert np.ptp(u1[IDN], axis=(0, 1)).max() == 0.0 and np.ptp(u2[IDN], axis=(0, 2)).max() == 0.0 \
            and np.ptp(u3[IDN], axis=(1, 2)).max() == 0.0      # planar waves stay planar
    finally:
        shutil.rmtree(wd, ignore_errors=True)


IDN, IEN = 0, 4
SYMMETRY_CASES = {"vl2_plm": ("vl2", 2, []), "rk3_ppm": ("rk3", 3, ["--nghost=3"])}


@pytest.mark.parametrize("case", sorted(SYMMETRY_CASES))
def test_axis_aligned_waves_are_symmetric(emul_binary, case):
    check_axis_symmetry(emul_binary, *SYMMETRY_CASES[case])


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(SYMMETRY_CASES))
def test_axis_aligned_waves_are_symmetric_on_gpu(cuda_lib, case):
    assert os.access(CUDA_BIN, os.X_OK), "athena_b200/bin/athena_b200 not built"
    check_axis_symmetry(CUDA_BIN, *SYMMETRY_CASES[case])


# Known answers the survey captured from the unmodified reference ("Oracle outputs
# captured in this survey"; g++ -O3): BASELINE configs[0] at its full size and an 8-block MHD wave.
# These do not need oracle/_ref at test time.
KNOWN_ANSWERS = {
    "hydro_64cubed_config0": dict(
        mhd=False, mesh=(64, 64, 64), block=(64, 64, 64), tlim=1.0, ncycle=143,
        errors={"RMS-L1-Error": "1.336316e-08", "d_L1": "6.501698e-09"},
        last_dt=3.1269611416973042e-03, hst_row=-1,
        hst={"mass": 6.7499999999971791e+00, "tot-E": 6.0749999999967788e+00}),
    "mhd_64x32x32_8blocks": dict(
        mhd=True, mesh=(64, 32, 32), block=(32, 16, 16), tlim=0.5, ncycle=72,
        errors={"RMS-L1-Error": "5.973017e-08"},
        first_dt=7.0317065028282708e-03, hst_row=0,
        hst={"mass": 6.7500000000002407e+00, "tot-E": 1.7043749999996553e+01, "1-ME": 3.9413321153603178e+00,
             "2-ME": 3.3816169015164186e+00, "3-ME": 3.6458009831252944e+00}),
}


def check_known_answer(binary, case):
    ka = KNOWN_ANSWERS[case]
    blocks = refrun.linwave_blocks(ka["mhd"], mesh=ka["mesh"], block=ka["block"], tlim=ka["tlim"])
    blocks["problem"]["compute_error"] = "true"
    wd, r = run_host(binary, blocks, "linear_wave", ka["mhd"])
    try:
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        head, row = open(os.path.join(wd, "linearwave-errors.dat")).read().splitlines()
        cols = dict(zip(head.lstrip("#").split(), row.split()))
        assert int(cols["Ncycle"]) == ka["ncycle"]
        for k, v in ka["errors"].items():
            assert cols[k] == v, (k, cols[k], v)
        dts = [float(l.split("dt=")[1]) for l in r.stdout.splitlines() if l.startswith("cycle=")]
        if "first_dt" in ka:
            assert dts[0] == ka["first_dt"]
        if "last_dt" in ka:
            assert dts[-1] == ka["last_dt"]
        names, h = refrun.read_hst(glob.glob(wd + "/*.hst")[0])
        for k, v in ka["hst"].items():
            got = h[ka["hst_row"], names.index(k)]
            assert abs(got - v) <= 1e-10 * abs(v), (k, got, v)   # north-star history tolerance
    finally:
        shutil.rmtree(wd, ignore_errors=True)


@pytest.mark.parametrize("case", sorted(KNOWN_ANSWERS))
def test_survey_known_answers(emul_binary, case):
    check_known_answer(emul_binary, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(KNOWN_ANSWERS))
def test_survey_known_answers_on_gpu(cuda_lib, case):
    assert os.access(CUDA_BIN, os.X_OK), "athena_b200/bin/athena_b200 not built"
    check_known_answer(CUDA_BIN, case)


# Checkpoint / resume (main.cpp -r, Mesh restart constructor, tst/regression/scripts/tests/outputs/
# restart tests): a run resumed from an intermediate .rst must end exactly where the uninterrupted
# run ends, and the two codes must be able to resume from each other's files.
def rst_equal(a, c):
    assert (a["ncycle"], a["time"], a["dt"], a["loc"]) == (c["ncycle"], c["time"], c["dt"], c["loc"])
    for k in ("u", "b1", "b2", "b3", "s"):
        assert len(a[k]) == len(c[k])
        for x, y in zip(a[k], c[k]):
            np.testing.assert_array_equal(x, y, err_msg=k)


def check_restart(binary, case):
    cfg, prob, mhd, mk = CASES[case][:4]
    extra = CASES[case][4] if len(CASES[case]) > 4 else []
    if not refrun.have_ref(cfg):
        pytest.skip("reference binary oracle/_ref/athena_%s not built" % cfg)
    blocks = mk()
    dirs = []
    try:
        wd_ref, _ = refrun.run_reference(cfg, blocks)
        wd, r = run_host(binary, blocks, prob, mhd, extra=extra)
        dirs += [wd_ref, wd]
        assert r.returncode == 0, r.stdout[-2000:]
        final_ref = refrun.read_rst(glob.glob(wd_ref + "/*.final.rst")[0])
        rst_equal(final_ref, refrun.read_rst(glob.glob(wd + "/*.final.rst")[0]))
        names, hst_full = refrun.read_hst(glob.glob(wd + "/*.hst")[0])
        flags = ["--prob=" + prob] + (["-b"] if mhd else []) + list(extra)
        # host resumes from its own checkpoint, and from the reference's
        for src in (wd, wd_ref):
            mid = sorted(glob.glob(src + "/*.00001.rst"))[0]
            wd2 = tempfile.mkdtemp(prefix="athhost_")
            dirs.append(wd2)
            r = subprocess.run([binary, "-r", mid, "-d", wd2] + flags, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            rst_equal(final_ref, refrun.read_rst(glob.glob(wd2 + "/*.final.rst")[0]))
            assert not glob.glob(wd2 + "/*.00000.rst") and not glob.glob(wd2 + "/*.00001.rst")   # numbering goes on
            _, hst_tail = refrun.read_hst(glob.glob(wd2 + "/*.hst")[0])
            np.testing.assert_array_equal(hst_tail, hst_full[-hst_tail.shape[0]:])
        # the reference resumes from the host's checkpoint
        mid = sorted(glob.glob(wd + "/*.00001.rst"))[0]
        wd3 = tempfile.mkdtemp(prefix="athref_")
        dirs.append(wd3)
        r = subprocess.run([refrun.ref_binary(cfg), "-r", mid, "-d", wd3], capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rst_equal(final_ref, refrun.read_rst(glob.glob(wd3 + "/*.final.rst")[0]))
        # -i next to -r overrides parameters of the checkpoint (main.cpp:240-245)
        wd4 = tempfile.mkdtemp(prefix="athhost_")
        dirs.append(wd4)
        r = subprocess.run([binary, "-r", mid, "-d", wd4, "time/nlim=%d" % (final_ref["ncycle"] - 1)] + flags,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "Terminating on cycle limit" in r.stdout
        # a changed output cadence after the restart: same dump schedule as the reference
        # (ParameterInput::RollbackNextTime / ForwardNextTime, parameter_input.cpp:754-835)
        new_dt = "output2/dt=%r" % (0.37 * blocks["output2"]["dt"])
        wd5, wd6 = tempfile.mkdtemp(prefix="athref_"), tempfile.mkdtemp(prefix="athhost_")
        dirs += [wd5, wd6]
        r = subprocess.run([refrun.ref_binary(cfg), "-r", mid, "-d", wd5, new_dt], capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert r.returncode == 0, r.stdout[-1500:]
        r = subprocess.run([binary, "-r", mid, "-d", wd6, new_dt] + flags, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:]
        names5 = sorted(os.path.basename(f) for f in glob.glob(wd5 + "/*.rst"))
        names6 = sorted(os.path.basename(f) for f in glob.glob(wd6 + "/*.rst"))
        assert names5 == names6 and len(names5) >= 2, (names5, names6)
        for nm in names5:
            rst_equal(refrun.read_rst(os.path.join(wd5, nm)), refrun.read_rst(os.path.join(wd6, nm)))
        # ... and so does an input file given with -i next to -r: only the blocks it names change
        over = os.path.join(wd4, "athinput.override")
        with open(over, "w") as f:
            f.write("<time>\ntlim = %r\n" % (0.5 * (final_ref["time"] + refrun.read_rst(mid)["time"])))
        r = subprocess.run([binary, "-r", mid, "-i", over, "-d", wd4] + flags, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0 and "Terminating on time limit" in r.stdout, r.stdout[-1500:]
        short = refrun.read_rst(glob.glob(wd4 + "/*.final.rst")[0])
        assert refrun.read_rst(mid)["time"] < short["time"] < final_ref["time"]
        assert short["ncycle"] < final_ref["ncycle"]
    finally:
        for d in dirs:
            shutil.rmtree(d, ignore_errors=True)


RESTART_CASES = ["mhd_blast", "kh_rk3_ppm_scalar", "briowu_x3_reflect_outflow", "smr_linwave_hydro_two_levels_rk3",
                 "smr_linwave_mhd_centre"]


@pytest.mark.parametrize("case", RESTART_CASES)
def test_restart_resumes_bitwise(emul_binary, case):
    check_restart(emul_binary, case)


def check_sr_restart(binary, case="sr_blast_hllc"):
    """Special relativity: a checkpoint holds U only, so after -r the fixed-point W(U) starts cold (w = w1 = 0,
    restart.cpp writes w / w1 under GENERAL_RELATIVITY only) and the reference's own restarted run is not its
    continuous run bit for bit.  What must hold: both codes, restarted from the same checkpoint, end identically."""
    cfg, prob, mhd, mk = CASES[case][:4]
    extra = CASES[case][4]
    if not refrun.have_ref(cfg):
        pytest.skip("reference binary oracle/_ref/athena_%s not built" % cfg)
    dirs = []
    try:
        wd_ref, _ = refrun.run_reference(cfg, mk())
        dirs.append(wd_ref)
        mid = sorted(glob.glob(wd_ref + "/*.00001.rst"))[0]
        wd_a, wd_b = tempfile.mkdtemp(prefix="athref_"), tempfile.mkdtemp(prefix="athhost_")
        dirs += [wd_a, wd_b]
        r = subprocess.run([refrun.ref_binary(cfg), "-r", mid, "-d", wd_a], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        r = subprocess.run([binary, "-r", mid, "-d", wd_b, "--prob=" + prob] + list(extra), capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rst_equal(refrun.read_rst(glob.glob(wd_a + "/*.final.rst")[0]), refrun.read_rst(glob.glob(wd_b + "/*.final.rst")[0]))
    finally:
        for d in dirs:
            shutil.rmtree(d, ignore_errors=True)


def test_sr_restart_matches_the_reference_restarted_from_the_same_checkpoint(emul_binary):
    check_sr_restart(emul_binary)


@pytest.mark.gpu
def test_sr_restart_matches_the_reference_restarted_from_the_same_checkpoint_on_gpu(cuda_lib):
    check_sr_restart(CUDA_BIN)


@pytest.mark.gpu
@pytest.mark.parametrize("case", RESTART_CASES)
def test_restart_resumes_bitwise_on_gpu(cuda_lib, case):
    assert os.access(CUDA_BIN, os.X_OK), "athena_b200/bin/athena_b200 not built"
    check_restart(CUDA_BIN, case)


def test_restart_with_a_disabled_output_block_and_output_block_checks(emul_binary):
    """outputs.cpp:206-223: `next_time` is added to every <outputN> block before the enable check, so
    a checkpoint written next to a disabled block (the common `dt = -1`) can be resumed; a block
    with neither dt nor dcycle, or with both, is a fatal error."""
    blocks = blast_blocks(n=32, blk=16, tlim=0.04)
    blocks["output3"] = {"file_type": "vtk", "variable": "prim", "dt": -1}
    dirs = []
    try:
        wd, r = run_host(emul_binary, blocks, "blast", True)
        dirs.append(wd)
        assert r.returncode == 0, r.stdout[-1500:]
        assert not glob.glob(wd + "/*.vtk")
        full = refrun.read_rst(glob.glob(wd + "/*.final.rst")[0])
        mid = sorted(glob.glob(wd + "/*.00001.rst"))[0]
        wd2 = tempfile.mkdtemp(prefix="athhost_")
        dirs.append(wd2)
        r = subprocess.run([emul_binary, "-r", mid, "-d", wd2, "--prob=blast", "-b"], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        rst_equal(full, refrun.read_rst(glob.glob(wd2 + "/*.final.rst")[0]))
        for bad, msg in (({"file_type": "vtk", "variable": "prim"}, "Either dt or dcycle must be specified"),
                         ({"file_type": "vtk", "variable": "prim", "dt": 0.01, "dcycle": 3},
                          "dt and dcycle cannot be specified simultaneously")):
            blocks["output3"] = bad
            wd3, r = run_host(emul_binary, blocks, "blast", True)
            dirs.append(wd3)
            assert r.returncode != 0 and msg in (r.stdout + r.stderr), (r.stdout + r.stderr)[-800:]
    finally:
        for d in dirs:
            shutil.rmtree(d, ignore_errors=True)


def test_wall_time_limit_and_signals_checkpoint_and_resume(emul_binary):
    """-t hh:mm:ss and SIGTERM (utils/signal_handler.cpp, main.cpp:568-618): the run stops after
    the current cycle, says why, and leaves a final restart dump; resuming from it ends on the
    bits of the uninterrupted run."""
    import signal
    import time
    blocks = blast_blocks(n=32, blk=16, tlim=0.12)
    blocks["output2"]["dt"] = 1.0      # no intermediate dumps: only the final one
    dirs = []
    try:
        wd_full, r = run_host(emul_binary, blocks, "blast", True)
        dirs.append(wd_full)
        assert r.returncode == 0 and "Terminating on time limit" in r.stdout
        full = refrun.read_rst(glob.glob(wd_full + "/*.final.rst")[0])
        for how in ("alarm", "sigterm"):
            wd = tempfile.mkdtemp(prefix="athhost_")
            dirs.append(wd)
            inp = os.path.join(wd, "athinput.host")
            with open(inp, "w") as f:
                f.write(refrun.format_athinput(blocks))
            args = [emul_binary, "-i", inp, "-d", wd, "--prob=blast", "-b"]
            if how == "alarm":
                r = subprocess.run(args + ["-t", "00:00:01"], capture_output=True, text=True, timeout=600)
                out, msg = r.stdout, "Terminating on wall-time limit"
            else:
                p = subprocess.Popen(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                time.sleep(1.0)
                p.send_signal(signal.SIGTERM)
                out, _ = p.communicate(timeout=600)
                msg = "Terminating on Terminate signal"
            if "Terminating on time limit" in out:
                pytest.skip("host finished the whole run within a second")
            assert msg in out, out[-1500:]
            part = refrun.read_rst(glob.glob(wd + "/*.final.rst")[0])
            assert 0 < part["ncycle"] < full["ncycle"] and part["time"] < full["time"]
            wd2 = tempfile.mkdtemp(prefix="athhost_")
            dirs.append(wd2)
            r = subprocess.run([emul_binary, "-r", glob.glob(wd + "/*.final.rst")[0], "-d", wd2, "--prob=blast", "-b"],
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0 and "Terminating on time limit" in r.stdout, r.stdout[-1500:]
            rst_equal(full, refrun.read_rst(glob.glob(wd2 + "/*.final.rst")[0]))
    finally:
        for d in dirs:
            shutil.rmtree(d, ignore_errors=True)


def test_parameter_input_error_behaviour(emul_binary):
    """Missing parameter / unknown cmdline override / unpaired periodic BC: message + exit status as
    ATHENA_ERROR does (src/defs.hpp.in:155-159, parameter_input.cpp:352-395)."""
    blocks = blast_blocks()
    bad = {k: dict(v) for k, v in blocks.items()}
    del bad["time"]["tlim"]
    wd, r = run_host(emul_binary, bad, "blast", True)
    assert r.returncode != 0 and "Parameter name 'tlim' not found in block 'time'" in r.stdout
    shutil.rmtree(wd, ignore_errors=True)
    wd, r = run_host(emul_binary, blocks, "blast", True, extra=["output1/new_param=3"])
    assert r.returncode != 0 and "on command line not found" in r.stdout
    shutil.rmtree(wd, ignore_errors=True)
    bad = {k: dict(v) for k, v in blocks.items()}
    bad["mesh"]["ix1_bc"] = "outflow"
    wd, r = run_host(emul_binary, bad, "blast", True)
    assert r.returncode != 0 and "both sides must be periodic" in r.stdout   # (ox1_bc is still periodic)
    shutil.rmtree(wd, ignore_errors=True)
    bad["mesh"]["ix1_bc"] = bad["mesh"]["ox1_bc"] = "user"   # ... and the blast problem enrolls no function
    wd, r = run_host(emul_binary, bad, "blast", True)
    assert r.returncode != 0 and "A user-defined boundary is specified but the actual boundary function" in r.stdout
    shutil.rmtree(wd, ignore_errors=True)
    bad["mesh"]["ix1_bc"] = bad["mesh"]["ox1_bc"] = "polar"
    wd, r = run_host(emul_binary, bad, "blast", True)
    assert r.returncode != 0 and "GetBoundaryFlag" in r.stdout
    shutil.rmtree(wd, ignore_errors=True)
    wd, r = run_host(emul_binary, blocks, "blast", True, extra=["time/tlim=0.004"])
    assert r.returncode == 0 and "Terminating on time limit" in r.stdout
    # restart files that do not fit the configuration are refused with the reference's diagnosis
    # (meshblock.cpp:411-417), never read into arrays of another shape
    rst = glob.glob(wd + "/*.final.rst")[0]
    for flags, text in ((["--prob=blast"], "input parameters are inconsistent"),
                        (["--prob=blast", "-b", "--nghost=3"], "input parameters are inconsistent"),
                        (["--prob=blast", "-b", "meshblock/nx1=8"], "The restart file holds 8 MeshBlocks")):
        r = subprocess.run([emul_binary, "-r", rst, "-d", wd] + flags, capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and text in r.stdout, r.stdout[-500:]
    r = subprocess.run([emul_binary, "-r", wd + "/missing.rst"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "could not be opened" in r.stdout
    shutil.rmtree(wd, ignore_errors=True)
