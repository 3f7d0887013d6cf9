"""TEST INFRASTRUCTURE (oracle harness) -- never imported by the product package.

Runs the UNMODIFIED reference binary built by ``oracle/Makefile.ref`` into
``oracle/_ref/athena_<cfg>`` and parses its full-precision artefacts:

* ``.rst`` restart dumps  (layout: /root/reference/src/outputs/restart.cpp:42-246)
* ``.hst`` history files  (layout: /root/reference/src/outputs/history.cpp:376-447)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm may
use this module.  It reads nothing under /root/reference at run time (the binary is
self-contained), so it works on the GPU box as well.
"""
import os
import re
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# cfg name -> make variables (see oracle/Makefile.ref)
CONFIGS = {
    "linwave_hydro_hllc": dict(PROB="linear_wave", MHD=0, FLUX="hllc"),
    "linwave_hydro_roe": dict(PROB="linear_wave", MHD=0, FLUX="roe"),
    "linwave_hydro_hlle": dict(PROB="linear_wave", MHD=0, FLUX="hlle"),
    "linwave_mhd_hlld": dict(PROB="linear_wave", MHD=1, FLUX="hlld"),
    "linwave_mhd_roe": dict(PROB="linear_wave", MHD=1, FLUX="roe"),
    "linwave_mhd_hlle": dict(PROB="linear_wave", MHD=1, FLUX="hlle"),
    "linwave_hydro_llf": dict(PROB="linear_wave", MHD=0, FLUX="llf"),
    "linwave_mhd_llf": dict(PROB="linear_wave", MHD=1, FLUX="llf"),
    "blast_mhd_hlld": dict(PROB="blast", MHD=1, FLUX="hlld"),
    "blast_mhd_roe": dict(PROB="blast", MHD=1, FLUX="roe"),
    "shocktube_hydro_roe": dict(PROB="shock_tube", MHD=0, FLUX="roe"),
    "shocktube_hydro_hllc": dict(PROB="shock_tube", MHD=0, FLUX="hllc"),
    "shocktube_mhd_hlld": dict(PROB="shock_tube", MHD=1, FLUX="hlld"),
    "shocktube_mhd_roe": dict(PROB="shock_tube", MHD=1, FLUX="roe"),
    "blast_hydro_hllc": dict(PROB="blast", MHD=0, FLUX="hllc"),
    "ot_mhd_hlld": dict(PROB="orszag_tang", MHD=1, FLUX="hlld"),
    "kh_hydro_hllc_s1g3": dict(PROB="kh", MHD=0, FLUX="hllc", NGHOST=3, NSCALARS=1),
    "kh_hydro_hllc_s2": dict(PROB="kh", MHD=0, FLUX="hllc", NSCALARS=2),
    "kh_mhd_hlld_s1": dict(PROB="kh", MHD=1, FLUX="hlld", NSCALARS=1),
    "linwave_hydro_hllc_g3": dict(PROB="linear_wave", MHD=0, FLUX="hllc", NGHOST=3),
    "linwave_mhd_hlld_g3": dict(PROB="linear_wave", MHD=1, FLUX="hlld", NGHOST=3),
    "cpaw_mhd_hlld": dict(PROB="cpaw", MHD=1, FLUX="hlld"),
    "fieldloop_mhd_hlld": dict(PROB="field_loop", MHD=1, FLUX="hlld"),
    "linwave_hydro_hllc_g4": dict(PROB="linear_wave", MHD=0, FLUX="hllc", NGHOST=4),   # PPM on a refined mesh
    "linwave_mhd_hlld_g4": dict(PROB="linear_wave", MHD=1, FLUX="hlld", NGHOST=4),     # ... with magnetic fields
    "usrsrc_hydro_hllc": dict(PROB="usrsrc", MHD=0, FLUX="hllc", PGEN="oracle"),   # oracle/pgen/usrsrc.cpp
    "usrsrc_mhd_hlld": dict(PROB="usrsrc", MHD=1, FLUX="hlld", PGEN="oracle"),
    "blast_iso_hydro_hlle": dict(PROB="blast", MHD=0, FLUX="hlle", EOS="isothermal"),
    "blast_iso_hydro_llf": dict(PROB="blast", MHD=0, FLUX="llf", EOS="isothermal"),
    "blast_iso_hydro_roe": dict(PROB="blast", MHD=0, FLUX="roe", EOS="isothermal"),
    "ot_iso_mhd_hlle": dict(PROB="orszag_tang", MHD=1, FLUX="hlle", EOS="isothermal"),
    "ot_iso_mhd_llf": dict(PROB="orszag_tang", MHD=1, FLUX="llf", EOS="isothermal"),
    "blast_sr_hydro_hlle": dict(PROB="blast", MHD=0, FLUX="hlle", SR=1),   # configure.py -s (SR part)
    "blast_sr_hydro_hllc": dict(PROB="blast", MHD=0, FLUX="hllc", SR=1),   # (-s default flux)
    "blast_sr_hydro_hllc_g3": dict(PROB="blast", MHD=0, FLUX="hllc", SR=1, NGHOST=3),   # PPM
    "blast_sr_hydro_llf": dict(PROB="blast", MHD=0, FLUX="llf", SR=1),
    "grtube_sr_hydro_hlle_s1": dict(PROB="gr_shock_tube", MHD=0, FLUX="hlle", SR=1, NSCALARS=1),
    "grlinwave_sr_hydro_hllc": dict(PROB="gr_linear_wave", MHD=0, FLUX="hllc", SR=1),   # inputs/hydro_sr/athinput.linear_wave
    "grtube_sr_hydro_hllc": dict(PROB="gr_shock_tube", MHD=0, FLUX="hllc", SR=1),   # inputs/hydro_sr/athinput.mb_*
    "ot_iso_mhd_roe": dict(PROB="orszag_tang", MHD=1, FLUX="roe", EOS="isothermal"),
    "ot_iso_mhd_hlld": dict(PROB="orszag_tang", MHD=1, FLUX="hlld", EOS="isothermal"),   # (blast.cpp:208 writes
    # u(IEN) unguarded with -b: the reference's own blast problem cannot run isothermal MHD)
    "kh_iso_mhd_hlld_s1": dict(PROB="kh", MHD=1, FLUX="hlld", EOS="isothermal", NSCALARS=1),
    "jet_hydro_hllc": dict(PROB="jet", MHD=0, FLUX="hllc"),       # user-enrolled inner-x1 boundary
    "jet_mhd_hlld": dict(PROB="jet", MHD=1, FLUX="hlld"),
}


def ref_binary(cfg):
    return os.path.join(REF_DIR, "athena_" + cfg)


def have_ref(cfg):
    return os.access(ref_binary(cfg), os.X_OK)


def build_ref(cfg, jobs=None, quiet=True):
    """Compile the reference for ``cfg`` (needs /root/reference; no-op if already built)."""
    if have_ref(cfg):
        return True
    if not os.path.isdir("/root/reference/src"):
        return False
    mk = CONFIGS[cfg]
    args = ["make", "-f", os.path.join(HERE, "Makefile.ref"), "CFG=" + cfg,
            "OUT=" + REF_DIR, "-j%d" % (jobs or os.cpu_count() or 4)]
    args += ["%s=%s" % (k, v) for k, v in mk.items()]
    r = subprocess.run(args, cwd=os.path.dirname(HERE), capture_output=quiet, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference build failed for %s:\n%s" % (cfg, (r.stderr or "")[-4000:]))
    # intermediates are not needed once linked
    subprocess.run(["rm", "-rf", os.path.join(REF_DIR, cfg)])
    return have_ref(cfg)


def format_athinput(blocks):
    """blocks: dict name -> dict(param -> value) in the reference's <block> grammar
    (/root/reference/src/parameter_input.cpp)."""
    out = []
    for bname, params in blocks.items():
        out.append("<%s>" % bname)
        for k, v in params.items():
            if isinstance(v, bool):
                v = "true" if v else "false"
            elif isinstance(v, float):
                v = repr(v)
            out.append("%-14s = %s" % (k, v))
        out.append("")
    return "\n".join(out)


def run_reference(cfg, blocks, workdir=None, threads=1, timeout=3600):
    """Run the reference binary on an athinput given as nested dicts. Returns (workdir, stdout)."""
    exe = ref_binary(cfg)
    if not have_ref(cfg):
        raise FileNotFoundError(exe + " (build it here with oracle/Makefile.ref)")
    workdir = workdir or tempfile.mkdtemp(prefix="athref_")
    os.makedirs(workdir, exist_ok=True)
    blocks = {k: dict(v) for k, v in blocks.items()}
    blocks.setdefault("mesh", {})["num_threads"] = threads
    inp = os.path.join(workdir, "athinput.ref")
    with open(inp, "w") as f:
        f.write(format_athinput(blocks))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    r = subprocess.run([exe, "-i", inp, "-d", workdir], capture_output=True, text=True,
                       timeout=timeout, env=env)
    if r.returncode != 0:
        raise RuntimeError("reference run failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return workdir, r.stdout


def parse_zone_cycles(stdout):
    """Extract the reference's own throughput line (/root/reference/src/main.cpp:639-651)."""
    out = {}
    for key in ("zone-cycles/cpu_second", "zone-cycles/omp_wsecond", "cpu time used",
                "omp wtime used", "zone-cycles"):
        m = re.search(re.escape(key) + r"\s*=\s*([0-9.eE+-]+)", stdout)
        if m:
            out[key] = float(m.group(1))
    m = re.search(r"cycle=(\d+)\s+time=", stdout[::-1][::-1].rsplit("cycle=", 1)[-1] and stdout)
    return out


def read_rst(path):
    """Parse a restart dump. Returns dict(params=str, nb, root_level, mesh_size, time, dt, ncycle,
    nx=(nx1,nx2,nx3) of the MeshBlock, ng, loc=[(lx1,lx2,lx3,level)], u=[...], b1=[...], ...)."""
    raw = open(path, "rb").read()
    end = raw.index(b"<par_end>\n") + len(b"<par_end>\n")
    params = raw[:end].decode()
    off = end
    nbtotal, root_level = struct.unpack_from("ii", raw, off)
    off += 8
    # RegionSize: 9 doubles (x1min,x2min,x3min,x1max,x2max,x3max,x1rat,x2rat,x3rat) + 3 ints + pad
    rs = struct.unpack_from("9d3i", raw, off)
    off += 88
    time, dt = struct.unpack_from("dd", raw, off)
    off += 16
    (ncycle,) = struct.unpack_from("i", raw, off)
    off += 4
    (datasize,) = struct.unpack_from("Q", raw, off)
    off += 8

    def par(block, name, default=None):
        m = re.search(r"<%s>(.*?)(?=\n<|\Z)" % block, params, re.S)
        if not m:
            return default
        mm = re.search(r"^\s*%s\s*=\s*([^#\n]+)" % name, m.group(1), re.M)
        return mm.group(1).strip() if mm else default

    mnx = [int(par("mesh", "nx%d" % d)) for d in (1, 2, 3)]
    bnx = [int(par("meshblock", "nx%d" % d, mnx[d - 1])) for d in (1, 2, 3)]
    locs = []
    for _ in range(nbtotal):
        lx1, lx2, lx3, lev = struct.unpack_from("qqqi", raw, off)
        off += 32
        off += 8  # cost
        locs.append((lx1, lx2, lx3, lev))
    # infer NGHOST, MHD and NSCALARS from datasize
    nd = datasize // 8
    found = None
    # (NHYDRO = 5, or 4 for the isothermal EOS: tried second, so an isothermal file with scalars reads
    #  as 5 + (nscalars - 1) components -- the same bytes in the same order)
    for nh in (5, 4):
        for ng in (2, 3, 4):
            nc = [n + 2 * ng if n > 1 else 1 for n in bnx]
            cells = nc[0] * nc[1] * nc[2]
            faces = ((nc[0] + 1) * nc[1] * nc[2] + nc[0] * (nc[1] + (1 if bnx[1] > 1 else 0)) * nc[2]
                     + nc[0] * nc[1] * (nc[2] + (1 if bnx[2] > 1 else 0)))
            for mhd in (0, 1):
                rest = nd - nh * cells - mhd * faces
                if rest >= 0 and rest % cells == 0 and rest // cells <= 8:
                    found = (ng, mhd, rest // cells, nc, nh)
                    break
            if found:
                break
        if found:
            break
    if not found:
        raise ValueError("cannot infer restart layout (datasize=%d)" % datasize)
    ng, mhd, nscal, nc, nh = found
    out = dict(params=params, nb=nbtotal, root_level=root_level, mesh_size=rs, time=time, dt=dt,
               ncycle=ncycle, nx=tuple(bnx), mesh_nx=tuple(mnx), ng=ng, mhd=bool(mhd),
               nscalars=nscal, loc=locs, u=[], b1=[], b2=[], b3=[], s=[])
    f2 = 1 if bnx[1] > 1 else 0
    f3 = 1 if bnx[2] > 1 else 0
    for _ in range(nbtotal):
        a = np.frombuffer(raw, dtype=np.float64, count=nd, offset=off)
        off += datasize
        p = 0
        n = nh * nc[2] * nc[1] * nc[0]
        out["u"].append(a[p:p + n].reshape(nh, nc[2], nc[1], nc[0]).copy())
        p += n
        if mhd:
            n = nc[2] * nc[1] * (nc[0] + 1)
            out["b1"].append(a[p:p + n].reshape(nc[2], nc[1], nc[0] + 1).copy())
            p += n
            n = nc[2] * (nc[1] + f2) * nc[0]
            out["b2"].append(a[p:p + n].reshape(nc[2], nc[1] + f2, nc[0]).copy())
            p += n
            n = (nc[2] + f3) * nc[1] * nc[0]
            out["b3"].append(a[p:p + n].reshape(nc[2] + f3, nc[1], nc[0]).copy())
            p += n
        if nscal:
            n = nscal * nc[2] * nc[1] * nc[0]
            out["s"].append(a[p:p + n].reshape(nscal, nc[2], nc[1], nc[0]).copy())
            p += n
    return out


def assemble(rst):
    """Re-assemble block interiors into global arrays. Returns dict(u, b1, b2, b3, s) where face
    arrays have the +1 extent of the global mesh."""
    ng = rst["ng"]
    nx1, nx2, nx3 = rst["nx"]
    M1, M2, M3 = rst["mesh_nx"]
    U = np.empty((5, M3, M2, M1))
    S = np.empty((rst["nscalars"], M3, M2, M1)) if rst["nscalars"] else None
    B1 = B2 = B3 = None
    if rst["mhd"]:
        B1 = np.empty((M3, M2, M1 + 1))
        B2 = np.empty((M3, M2 + 1, M1))
        B3 = np.empty((M3 + 1, M2, M1))
    g1 = ng
    g2 = ng if nx2 > 1 else 0
    g3 = ng if nx3 > 1 else 0
    for b, (lx1, lx2, lx3, _) in enumerate(rst["loc"]):
        o1, o2, o3 = lx1 * nx1, lx2 * nx2, lx3 * nx3
        U[:, o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1] = \
            rst["u"][b][:, g3:g3 + nx3, g2:g2 + nx2, g1:g1 + nx1]
        if S is not None:
            S[:, o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1] = \
                rst["s"][b][:, g3:g3 + nx3, g2:g2 + nx2, g1:g1 + nx1]
        if rst["mhd"]:
            B1[o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1 + 1] = \
                rst["b1"][b][g3:g3 + nx3, g2:g2 + nx2, g1:g1 + nx1 + 1]
            B2[o3:o3 + nx3, o2:o2 + nx2 + 1, o1:o1 + nx1] = \
                rst["b2"][b][g3:g3 + nx3, g2:g2 + nx2 + 1, g1:g1 + nx1]
            B3[o3:o3 + nx3 + 1, o2:o2 + nx2, o1:o1 + nx1] = \
                rst["b3"][b][g3:g3 + nx3 + 1, g2:g2 + nx2, g1:g1 + nx1]
    return dict(u=U, b1=B1, b2=B2, b3=B3, s=S)


def read_hst(path):
    """History file -> (column names, 2-D array)."""
    names = None
    rows = []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                if "[1]" in line:
                    names = re.findall(r"\[\d+\]=(\S+)", line)
                continue
            if line.strip():
                rows.append([float(x) for x in line.split()])
    return names, np.array(rows)


def kh_blocks(mesh=(32, 32, 8), block=None, tlim=0.05, integrator="rk3", xorder=3, cfl=0.3,
              rst_dt=None, amp=0.01, vflow=1.0, drho_rho0=0.0, iprob=4, b0=None):
    """athinput for pgen/kh.cpp iprob=4 (Lecoanet et al. 2016: periodic double shear layer with a
    passive-scalar dye), extruded to 3-D; inputs/hydro/athinput.kh-shear-lecoanet without the
    diffusion terms."""
    block = block or mesh
    extra = {} if b0 is None else {"b0": b0}
    return {
        "job": {"problem_id": "kh"},
        "output1": {"file_type": "hst", "dt": 1.0e-9 if rst_dt is None else rst_dt,
                    "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 1.0e30 if rst_dt is None else rst_dt},
        "time": {"cfl_number": cfl, "nlim": -1, "tlim": tlim, "integrator": integrator,
                 "xorder": xorder, "ncycle_out": 1},
        "mesh": {"nx1": mesh[0], "x1min": -0.5, "x1max": 0.5, "ix1_bc": "periodic",
                 "ox1_bc": "periodic",
                 "nx2": mesh[1], "x2min": -1.0, "x2max": 1.0, "ix2_bc": "periodic",
                 "ox2_bc": "periodic",
                 "nx3": mesh[2], "x3min": -0.25, "x3max": 0.25, "ix3_bc": "periodic",
                 "ox3_bc": "periodic", "num_threads": 1, "refinement": "none"},
        "meshblock": {"nx1": block[0], "nx2": block[1], "nx3": block[2]},
        "hydro": {"gamma": 1.6666667, "iso_sound_speed": 1.0},
        "problem": dict({"iprob": iprob, "amp": amp, "vflow": vflow, "vboost": 0.0,
                         "drho_rho0": drho_rho0}, **extra),
    }


BC_KEYS = ("ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc", "ox3_bc")


def set_bcs(blocks, bcs):
    """bcs: six <mesh> boundary flags ('periodic' | 'reflecting' | 'outflow') or None."""
    if bcs is not None:
        for key, val in zip(BC_KEYS, bcs):
            blocks["mesh"][key] = val
    return blocks


def blast_blocks(mesh=(32, 32, 32), block=(16, 16, 16), tlim=0.02, integrator="vl2", xorder=2, cfl=0.3,
                 rst_dt=None, bcs=None, center=(0.0, 0.0, 0.0), prat=100.0, b0=1.0, radius=0.1):
    """athinput for pgen/blast.cpp in the box [-1,1]^3 (inputs/mhd/athinput.blast shape)."""
    b = {
        "job": {"problem_id": "Blast"},
        "output1": {"file_type": "hst", "dt": 1.0e-9 if rst_dt is None else rst_dt,
                    "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 1.0e30 if rst_dt is None else rst_dt},
        "time": {"cfl_number": cfl, "nlim": -1, "tlim": tlim, "integrator": integrator, "xorder": xorder,
                 "ncycle_out": 1},
        "mesh": {"nx1": mesh[0], "x1min": -1.0, "x1max": 1.0, "ix1_bc": "periodic", "ox1_bc": "periodic",
                 "nx2": mesh[1], "x2min": -1.0, "x2max": 1.0, "ix2_bc": "periodic", "ox2_bc": "periodic",
                 "nx3": mesh[2], "x3min": -1.0, "x3max": 1.0, "ix3_bc": "periodic", "ox3_bc": "periodic",
                 "num_threads": 1},
        "meshblock": {"nx1": block[0], "nx2": block[1], "nx3": block[2]},
        "hydro": {"gamma": 1.666666666667, "iso_sound_speed": 0.4082482905},
        "problem": {"pamb": 0.1, "prat": prat, "radius": radius, "damb": 1.0, "drat": 1.0,
                    "ramp": 0.0, "b0": b0, "angle": 30, "x1_0": center[0], "x2_0": center[1],
                    "x3_0": center[2]},
    }
    return set_bcs(b, bcs)


def ot_blocks(mesh=(32, 32, 8), block=(16, 16, 8), tlim=0.05, integrator="vl2", xorder=2, cfl=0.3, rst_dt=None):
    """athinput for pgen/orszag_tang.cpp (inputs/mhd/athinput.orszag-tang), extruded along x3."""
    b = kh_blocks(mesh=mesh, block=block, tlim=tlim, integrator=integrator, xorder=xorder, cfl=cfl, rst_dt=rst_dt)
    b["job"]["problem_id"] = "OrszagTang"
    b["mesh"].update(x1min=-0.5, x1max=0.5, x2min=-0.5, x2max=0.5)
    b["hydro"] = {"gamma": 1.666666667, "iso_sound_speed": 1.0}
    b["problem"] = {}
    return b


def shocktube_blocks(mhd, left, right, shock_dir=1, xshock=None, bx=0.0, bcs=None, **kw):
    """athinput for pgen/shock_tube.cpp on the linear-wave test box (periodic, so the two
    discontinuities interact across the seam).  left/right = (d, u, v, w, p[, by, bz]) in the
    shock_dir frame (shock_tube.cpp:335-366)."""
    b = linwave_blocks(mhd, problem_id="Sod", **kw)
    ext = {1: 3.0, 2: 1.5, 3: 1.5}[shock_dir]
    pr = {"shock_dir": shock_dir, "xshock": 0.5 * ext if xshock is None else xshock,
          "compute_error": "false"}
    for side, st in (("l", left), ("r", right)):
        for key, val in zip(("d", "u", "v", "w", "p"), st[:5]):
            pr[key + side] = float(val)
        if mhd:
            pr["bx" + side], pr["by" + side], pr["bz" + side] = float(bx), float(st[5]), float(st[6])
    b["problem"] = pr
    return set_bcs(b, bcs)


def periodic_mesh(mesh, block, lo, hi):
    m = {}
    for d in range(3):
        m.update({"nx%d" % (d + 1): mesh[d], "x%dmin" % (d + 1): lo[d], "x%dmax" % (d + 1): hi[d],
                  "ix%d_bc" % (d + 1): "periodic", "ox%d_bc" % (d + 1): "periodic"})
    m.update(num_threads=1, refinement="none")
    block = block or mesh
    return m, {"nx1": block[0], "nx2": block[1], "nx3": block[2]}


def cpaw_blocks(mesh=(32, 16, 16), block=(16, 8, 8), tlim=1.0, rst_dt=None, integrator="vl2", xorder=2,
                cfl=0.3, v_par=0.0, direction=1, compute_error=True):
    """Circularly polarised Alfven wave (inputs/mhd/athinput.cpaw2d parameters on the 3-D box of the
    linear-wave inputs)."""
    m, mb = periodic_mesh(mesh, block, (0.0, 0.0, 0.0), (3.0, 1.5, 1.5))
    return {
        "job": {"problem_id": "cpaw"},
        "output1": {"file_type": "hst", "dt": 0.05, "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 1.0e30 if rst_dt is None else rst_dt},
        "time": {"cfl_number": cfl, "nlim": -1, "tlim": tlim, "integrator": integrator, "xorder": xorder,
                 "ncycle_out": 1},
        "mesh": m, "meshblock": mb,
        "hydro": {"gamma": 1.666666666666667, "iso_sound_speed": 1.0},
        "problem": {"compute_error": "true" if compute_error else "false", "b_par": 1.0, "b_perp": 0.1,
                    "pres": 0.1, "v_par": v_par, "dir": direction},
    }


def fieldloop_blocks(iprob=1, mesh=(32, 16, 16), block=(16, 8, 8), tlim=0.3, rst_dt=None, integrator="vl2",
                     xorder=2, cfl=0.3, drat=1.0):
    """Field-loop advection (inputs/mhd/athinput.field_loop parameters, 3-D box)."""
    m, mb = periodic_mesh(mesh, block, (-1.0, -0.5, -0.5), (1.0, 0.5, 0.5))
    return {
        "job": {"problem_id": "Loop"},
        "output1": {"file_type": "hst", "dt": 0.05, "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 1.0e30 if rst_dt is None else rst_dt},
        "time": {"cfl_number": cfl, "nlim": -1, "tlim": tlim, "integrator": integrator, "xorder": xorder,
                 "ncycle_out": 1},
        "mesh": m, "meshblock": mb,
        "hydro": {"iso_sound_speed": 1.732, "gamma": 1.66667},
        "problem": {"rad": 0.3, "amp": 1.0e-3, "vflow": 3.0, "iprob": iprob, "drat": drat},
    }


def linwave_blocks(mhd, mesh=(32, 16, 16), block=None, tlim=0.1, nlim=-1, wave_flag=0,
                   integrator="vl2", xorder=2, cfl=0.3, amp=1.0e-6, rst_dt=None,
                   vflow=0.0, problem_id="LinWave"):
    """athinput for the linear-wave problem (inputs/{hydro,mhd}/athinput.linear_wave3d shape)."""
    block = block or mesh
    b = {
        "job": {"problem_id": problem_id},
        "output1": {"file_type": "hst", "dt": 1.0e-9 if rst_dt is None else rst_dt,
                    "data_format": "%24.16e"},
        "output2": {"file_type": "rst", "dt": 1.0e30 if rst_dt is None else rst_dt},
        "time": {"cfl_number": cfl, "nlim": nlim, "tlim": tlim, "integrator": integrator,
                 "xorder": xorder, "ncycle_out": 1},
        "mesh": {"nx1": mesh[0], "x1min": 0.0, "x1max": 3.0, "ix1_bc": "periodic",
                 "ox1_bc": "periodic",
                 "nx2": mesh[1], "x2min": 0.0, "x2max": 1.5, "ix2_bc": "periodic",
                 "ox2_bc": "periodic",
                 "nx3": mesh[2], "x3min": 0.0, "x3max": 1.5, "ix3_bc": "periodic",
                 "ox3_bc": "periodic", "num_threads": 1, "refinement": "none"},
        "meshblock": {"nx1": block[0], "nx2": block[1], "nx3": block[2]},
        "hydro": {"gamma": 1.666666666666667, "iso_sound_speed": 1.0},
        "problem": {"compute_error": "false", "wave_flag": wave_flag, "amp": amp,
                    "vflow": vflow, "ang_2_vert": "false", "ang_3_vert": "false"},
    }
    return b
