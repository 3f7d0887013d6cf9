"""Generates the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/athena_<cfg>, built from /root/reference by oracle/Makefile.ref).  The reference has
no golden vectors for Newtonian hydro/MHD, so its own outputs are the
vectors.  Run from the repo root:  python tests/golden/make_golden.py [case ...]

Each fixture is an .npz holding the initial state (reference pgen output, cycle 0 restart dump),
the final state (final restart dump), both re-assembled to global arrays, the .hst rows printed
with %24.16e, and the run parameters.
"""
import glob
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # BASELINE configs[1] shape at parity size: MHD VL2+PLM+HLLD+CT, 8 MeshBlocks
    "mhd_linwave_vl2_plm_hlld": dict(cfg="linwave_mhd_hlld", mhd=True, mesh=(32, 16, 16),
                                     block=(16, 8, 8), tlim=0.1, integrator="vl2", xorder=2),
    # BASELINE configs[0] shape at parity size: hydro VL2+PLM+HLLC
    "hydro_linwave_vl2_plm_hllc": dict(cfg="linwave_hydro_hllc", mhd=False, mesh=(32, 16, 16),
                                       block=(16, 8, 8), tlim=0.1, integrator="vl2", xorder=2),
    "mhd_linwave_rk2_plm_hlld": dict(cfg="linwave_mhd_hlld", mhd=True, mesh=(32, 16, 16),
                                     block=(16, 16, 16), tlim=0.05, integrator="rk2", xorder=2),
    "mhd_linwave_rk3_plm_hlld": dict(cfg="linwave_mhd_hlld", mhd=True, mesh=(32, 16, 16),
                                     block=(32, 16, 16), tlim=0.05, integrator="rk3", xorder=2),
    "hydro_linwave_rk3_plm_hllc": dict(cfg="linwave_hydro_hllc", mhd=False, mesh=(32, 16, 16),
                                       block=(16, 8, 8), tlim=0.1, integrator="rk3", xorder=2,
                                       wave_flag=4, vflow=0.5),
    # PPM (xorder=3) needs NGHOST >= 3: reference built with --nghost=3
    "hydro_linwave_rk3_ppm_hllc": dict(cfg="linwave_hydro_hllc_g3", mhd=False, mesh=(32, 16, 16),
                                       block=(16, 8, 8), tlim=0.1, integrator="rk3", xorder=3),
    "mhd_linwave_vl2_ppm_hlld": dict(cfg="linwave_mhd_hlld_g3", mhd=True, mesh=(32, 16, 16),
                                     block=(16, 8, 8), tlim=0.05, integrator="vl2", xorder=3,
                                     wave_flag=2, amp=1.0e-3),
    # the alternative Riemann solvers (one reference build per --flux)
    "hydro_linwave_vl2_plm_roe": dict(cfg="linwave_hydro_roe", mhd=False, mesh=(32, 16, 16),
                                      block=(16, 8, 8), tlim=0.1, integrator="vl2", xorder=2,
                                      amp=1.0e-3, flux="roe"),
    # supersonic background flow: the upwind overrides of roe.cpp:133-146 / bm = 0 of hlle.cpp
    "hydro_linwave_rk2_plm_roe_supersonic": dict(cfg="linwave_hydro_roe", mhd=False, mesh=(32, 16, 16),
                                                 block=(16, 16, 16), tlim=0.05, integrator="rk2",
                                                 xorder=2, amp=1.0e-2, vflow=2.0, flux="roe"),
    "hydro_linwave_vl2_plm_hlle": dict(cfg="linwave_hydro_hlle", mhd=False, mesh=(32, 16, 16),
                                       block=(16, 8, 8), tlim=0.1, integrator="vl2", xorder=2,
                                       amp=1.0e-3, vflow=0.7, flux="hlle"),
    "hydro_linwave_vl2_plm_llf": dict(cfg="linwave_hydro_llf", mhd=False, mesh=(32, 16, 16),
                                      block=(32, 16, 16), tlim=0.05, integrator="vl2", xorder=2,
                                      amp=1.0e-3, flux="llf"),
    "mhd_linwave_vl2_plm_roe": dict(cfg="linwave_mhd_roe", mhd=True, mesh=(32, 16, 16),
                                    block=(16, 8, 8), tlim=0.05, integrator="vl2", xorder=2,
                                    amp=1.0e-3, flux="roe"),
    "mhd_linwave_rk3_plm_roe_slow": dict(cfg="linwave_mhd_roe", mhd=True, mesh=(32, 16, 16),
                                         block=(16, 16, 16), tlim=0.03, integrator="rk3", xorder=2,
                                         wave_flag=2, amp=1.0e-2, vflow=1.0, flux="roe"),
    "mhd_linwave_vl2_plm_hlle": dict(cfg="linwave_mhd_hlle", mhd=True, mesh=(32, 16, 16),
                                     block=(16, 8, 8), tlim=0.05, integrator="vl2", xorder=2,
                                     amp=1.0e-3, flux="hlle"),
    "mhd_linwave_vl2_plm_llf": dict(cfg="linwave_mhd_llf", mhd=True, mesh=(32, 16, 16),
                                    block=(32, 16, 16), tlim=0.03, integrator="vl2", xorder=2,
                                    amp=1.0e-3, flux="llf"),
    # strong double rarefaction (Einfeldt 1-2-3) + the collision across the periodic seam: Roe's
    # intermediate densities go negative -> LLF fallback (roe.cpp:150-164, roe_mhd.cpp:189-205)
    "hydro_shocktube_vl2_plm_roe_llf": dict(cfg="shocktube_hydro_roe", mhd=False, mesh=(32, 16, 16),
                                            block=(16, 8, 8), tlim=0.05, integrator="vl2", xorder=2,
                                            flux="roe", shocktube=dict(
                                                left=(1.0, -2.0, 0.0, 0.0, 0.4),
                                                right=(1.0, 2.0, 0.0, 0.0, 0.4), shock_dir=1)),
    "mhd_shocktube_vl2_plm_roe_llf": dict(cfg="shocktube_mhd_roe", mhd=True, mesh=(32, 16, 16),
                                          block=(16, 8, 8), tlim=0.03, integrator="vl2", xorder=2,
                                          flux="roe", shocktube=dict(
                                              left=(1.0, -2.0, 0.3, 0.0, 0.4, 1.0, 0.2),
                                              right=(0.5, 2.0, -0.3, 0.1, 0.3, -1.0, 0.0),
                                              shock_dir=2, bx=0.75)),
    # SURVEY 8 row a10 / BASELINE configs[3] at parity size: KH (Lecoanet) + passive-scalar dye
    "hydro_kh_rk3_ppm_hllc_scalar": dict(cfg="kh_hydro_hllc_s1g3", mhd=False, mesh=(32, 32, 8),
                                         block=(16, 16, 8), tlim=0.03, integrator="rk3", xorder=3,
                                         nscalars=1, kh=dict(drho_rho0=0.5)),
    "hydro_kh_vl2_plm_hllc_2scalars": dict(cfg="kh_hydro_hllc_s2", mhd=False, mesh=(32, 32, 8),
                                           block=(16, 16, 8), tlim=0.03, integrator="vl2", xorder=2,
                                           nscalars=2, kh=dict()),
    "mhd_kh_rk2_plm_hlld_scalar": dict(cfg="kh_mhd_hlld_s1", mhd=True, mesh=(32, 32, 8),
                                       block=(16, 16, 8), tlim=0.03, integrator="rk2", xorder=2,
                                       nscalars=1, kh=dict(b0=0.5)),
    # physical boundary conditions (outflow / reflecting), incl. edges and corners
    "hydro_sod_vl2_plm_hllc_outflow_x1": dict(cfg="shocktube_hydro_hllc", mhd=False, mesh=(32, 16, 16),
                                              block=(16, 8, 8), tlim=0.25, integrator="vl2", xorder=2,
                                              bcs=["outflow", "outflow"] + ["periodic"] * 4,
                                              shocktube=dict(left=(1.0, 0.0, 0.0, 0.0, 1.0),
                                                             right=(0.125, 0.0, 0.0, 0.0, 0.1), shock_dir=1)),
    "mhd_briowu_rk2_plm_hlld_outflow_x2": dict(cfg="shocktube_mhd_hlld", mhd=True, mesh=(32, 16, 16),
                                               block=(16, 8, 8), tlim=0.06, integrator="rk2", xorder=2,
                                               bcs=["periodic"] * 2 + ["outflow"] * 2 + ["periodic"] * 2,
                                               shocktube=dict(left=(1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0),
                                                              right=(0.125, 0.0, 0.0, 0.0, 0.1, -1.0, 0.0),
                                                              shock_dir=2, bx=0.75)),
    "mhd_blast_vl2_plm_hlld_reflecting_box": dict(cfg="blast_mhd_hlld", mhd=True, mesh=(32, 32, 32),
                                                  block=(16, 16, 16), tlim=0.1, integrator="vl2", xorder=2,
                                                  bcs=["reflecting"] * 6,
                                                  blast=dict(center=(0.6, 0.55, 0.5), radius=0.3)),
    "hydro_blast_rk3_plm_hllc_mixed_bcs": dict(cfg="blast_hydro_hllc", mhd=False, mesh=(32, 32, 32),
                                               block=(16, 16, 16), tlim=0.1, integrator="rk3", xorder=2,
                                               bcs=["outflow", "reflecting", "reflecting", "outflow",
                                                    "periodic", "periodic"],
                                               blast=dict(center=(0.55, -0.6, 0.0), radius=0.3)),
    # extents that are neither powers of two nor multiples of 4 (row padding, tiles straddling rows)
    "mhd_linwave_vl2_plm_hlld_oddmesh": dict(cfg="linwave_mhd_hlld", mhd=True, mesh=(24, 20, 28),
                                             block=(12, 10, 14), tlim=0.05, integrator="vl2", xorder=2,
                                             amp=1.0e-3),
    "hydro_linwave_rk3_ppm_hllc_oddmesh": dict(cfg="linwave_hydro_hllc_g3", mhd=False, mesh=(22, 18, 26),
                                               block=(11, 9, 13), tlim=0.05, integrator="rk3", xorder=3,
                                               amp=1.0e-3, wave_flag=1, vflow=0.3),
    # vector-potential problems of the reference's MHD regression set: a nonlinear exact solution
    # (cpaw) and a loop in an otherwise field-free box (HLLD / CT with B == 0 almost everywhere)
    "mhd_cpaw_rk3_plm_hlld": dict(cfg="cpaw_mhd_hlld", mhd=True, mesh=(32, 16, 16), block=(16, 8, 8),
                                  tlim=0.1, integrator="rk3", xorder=2, cpaw=dict(v_par=1.0, direction=2)),
    "mhd_fieldloop_vl2_plm_hlld": dict(cfg="fieldloop_mhd_hlld", mhd=True, mesh=(32, 16, 16), block=(16, 8, 8),
                                       tlim=0.1, integrator="vl2", xorder=2, fieldloop=dict(iprob=5, drat=2.0)),
    # SRC_TERM: constant accelerations along all three axes (hydro/srcterms/constant_acc.cpp)
    "mhd_blast_rk2_plm_hlld_gravity": dict(cfg="blast_mhd_hlld", mhd=True, mesh=(32, 32, 32), block=(16, 16, 16),
                                           tlim=0.05, integrator="rk2", xorder=2,
                                           bcs=["periodic"] * 4 + ["reflecting"] * 2,
                                           blast=dict(center=(0.1, 0.0, -0.2), radius=0.3),
                                           grav_acc=(0.3, -0.2, -0.7)),
    "mhd_linwave_alfven_vl2": dict(cfg="linwave_mhd_hlld", mhd=True, mesh=(32, 16, 16),
                                   block=(8, 8, 8), tlim=0.05, integrator="vl2", xorder=2,
                                   wave_flag=1, amp=1.0e-2),
}


def make(name):
    c = CASES[name]
    if not refrun.build_ref(c["cfg"]):
        raise SystemExit("reference binary for %s unavailable" % c["cfg"])
    xmin, xmax, gamma = [0.0, 0.0, 0.0], [3.0, 1.5, 1.5], 1.666666666666667
    if "kh" in c:
        blocks = refrun.kh_blocks(mesh=c["mesh"], block=c["block"], tlim=c["tlim"],
                                  integrator=c["integrator"], xorder=c["xorder"], **c["kh"])
        m = blocks["mesh"]
        xmin, xmax = [m["x1min"], m["x2min"], m["x3min"]], [m["x1max"], m["x2max"], m["x3max"]]
        gamma = blocks["hydro"]["gamma"]
    elif "blast" in c:
        blocks = refrun.blast_blocks(mesh=c["mesh"], block=c["block"], tlim=c["tlim"],
                                     integrator=c["integrator"], xorder=c["xorder"], bcs=c.get("bcs"),
                                     **c["blast"])
        xmin, xmax, gamma = [-1.0] * 3, [1.0] * 3, blocks["hydro"]["gamma"]
    elif "cpaw" in c:
        blocks = refrun.cpaw_blocks(mesh=c["mesh"], block=c["block"], tlim=c["tlim"], integrator=c["integrator"],
                                    xorder=c["xorder"], compute_error=False, **c["cpaw"])
        blocks["output1"]["dt"] = 1.0e-9      # a history row every cycle (the dt sequence)
    elif "fieldloop" in c:
        blocks = refrun.fieldloop_blocks(mesh=c["mesh"], block=c["block"], tlim=c["tlim"],
                                         integrator=c["integrator"], xorder=c["xorder"], **c["fieldloop"])
        blocks["output1"]["dt"] = 1.0e-9
        xmin, xmax, gamma = [-1.0, -0.5, -0.5], [1.0, 0.5, 0.5], blocks["hydro"]["gamma"]
    elif "shocktube" in c:
        blocks = refrun.shocktube_blocks(c["mhd"], mesh=c["mesh"], block=c["block"], tlim=c["tlim"],
                                         integrator=c["integrator"], xorder=c["xorder"], bcs=c.get("bcs"),
                                         **c["shocktube"])
    else:
        blocks = refrun.linwave_blocks(c["mhd"], mesh=c["mesh"], block=c["block"], tlim=c["tlim"],
                                       integrator=c["integrator"], xorder=c["xorder"],
                                       wave_flag=c.get("wave_flag", 0), amp=c.get("amp", 1.0e-6),
                                       vflow=c.get("vflow", 0.0))
    for d, gd in enumerate(c.get("grav_acc", ())):
        blocks["hydro"]["grav_acc%d" % (d + 1)] = gd
    wd, out = refrun.run_reference(c["cfg"], blocks)
    r0 = refrun.read_rst(glob.glob(wd + "/*.00000.rst")[0])
    rf = refrun.read_rst(glob.glob(wd + "/*.final.rst")[0])
    a0, af = refrun.assemble(r0), refrun.assemble(rf)
    names, hst = refrun.read_hst(glob.glob(wd + "/*.hst")[0])
    meta = dict(c, name=name, ng=r0["ng"], ncycle=rf["ncycle"], time=rf["time"], dt_final=rf["dt"],
                dt0=r0["dt"], xmin=xmin, xmax=xmax, gamma=gamma,
                cfl=0.3, loc=[list(l) for l in r0["loc"]], hst_names=names,
                reference="Athena++ v24.0, g++ -O3 -std=c++11 -fopenmp, oracle/Makefile.ref")
    arrs = dict(meta=np.array(json.dumps(meta)), hst=hst, u0=a0["u"], uf=af["u"])
    if c["mhd"]:
        for k in ("b1", "b2", "b3"):
            arrs[k + "_0"] = a0[k]
            arrs[k + "_f"] = af[k]
    if c.get("nscalars"):
        arrs["s0"], arrs["sf"] = a0["s"], af["s"]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    shutil.rmtree(wd, ignore_errors=True)
    print("wrote", name, "cycles", rf["ncycle"], "size",
          os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    for n in (sys.argv[1:] or CASES):
        make(n)
