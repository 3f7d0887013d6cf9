"""The C++ host binary (Mesh / MeshBlock / TaskList / pgen API, athena_b200/bin/athena_b200) on the
bench workload and on the other BASELINE.json configuration shapes.  Prints one JSON line with the
binary's own zone-cycles/gpu_wsecond (main-loop wall clock, as the reference prints
zone-cycles/wsecond).
usage: python tools/host_bench.py [--config blast|config0|config1|config3] [--gpus N] [--cycles K]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def format_athinput(blocks):
    out = []
    for bname, params in blocks.items():
        out.append("<%s>" % bname)
        for k, v in params.items():
            out.append("%-14s = %s" % (k, repr(v) if isinstance(v, float) else v))
        out.append("")
    return "\n".join(out)


def mesh_block(mesh, lo, hi, block, **extra):
    m = {}
    for d, n, a, b in zip("123", mesh, lo, hi):
        m.update({"nx" + d: n, "x%smin" % d: a, "x%smax" % d: b, "ix%s_bc" % d: "periodic", "ox%s_bc" % d: "periodic"})
    m.update(extra)
    return {"mesh": m, "meshblock": {"nx1": block[0], "nx2": block[1], "nx3": block[2]}}


def base(problem_id, integrator, xorder, tlim, nlim):
    return {"job": {"problem_id": problem_id},
            "output1": {"file_type": "hst", "dt": 1.0e30},
            "time": {"cfl_number": 0.3, "nlim": nlim, "tlim": tlim, "integrator": integrator, "xorder": xorder,
                     "ncycle_out": 50}}


def run_binary(blocks, flags, label, zones, gpus):
    wd = tempfile.mkdtemp(prefix="athhostbench_")
    inp = os.path.join(wd, "athinput")
    open(inp, "w").write(format_athinput(blocks))
    exe = os.path.join(ROOT, "athena_b200", "bin", "athena_b200")
    r = subprocess.run([exe, "-i", inp, "-d", wd] + flags, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:], file=sys.stderr)
        sys.exit(1)
    zc = float(re.search(r"zone-cycles/gpu_wsecond = (\S+)", r.stdout).group(1))
    wt = float(re.search(r"gpu wtime used = (\S+)", r.stdout).group(1))
    ncyc = int(re.search(r"cycle=(\d+)\ntlim", r.stdout).group(1))
    print(json.dumps({"host": "athena_b200/bin/athena_b200 (C++ Mesh/TaskList host, one process, <mesh> num_gpus=%d)" % gpus,
                      "workload": label, "metric": "zone-cycles/gpu_wsecond (main loop wall clock)", "value": zc,
                      "n_gpus": gpus, "zones": zones, "cycles": ncyc, "main_loop_s": wt,
                      "ms_per_cycle": 1e3 * wt / max(ncyc, 1)}))
    subprocess.run(["rm", "-rf", wd])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cycles", type=int, default=30)
    ap.add_argument("--per-gpu", type=int, default=256)
    ap.add_argument("--block", type=int, default=64)
    ap.add_argument("--config", default="blast", choices=["blast", "config0", "config1", "config3", "smr", "sr"],
                    help="blast: the bench workload; config0/1/3: BASELINE.json configs[0], [1], [3] shapes; "
                         "smr: the MHD blast on a statically refined mesh (operator sequence + box lists); "
                         "sr: special-relativistic hydro blast (-s, HLLC)")
    a = ap.parse_args()
    if a.config == "config0":   # 3-D hydro linear wave 64^3, VL2+PLM+HLLC, single MeshBlock, tlim=1
        b = base("LinWave", "vl2", 2, 1.0, -1)
        b.update(mesh_block((64, 64, 64), (0.0, 0.0, 0.0), (3.0, 1.5, 1.5), (64, 64, 64)))
        b["hydro"] = {"gamma": 1.666666666666667, "iso_sound_speed": 1.0}
        b["problem"] = {"compute_error": "false", "wave_flag": 0, "amp": 1.0e-6, "vflow": 0.0}
        return run_binary(b, ["--prob=linear_wave"], "BASELINE configs[0]: hydro linear wave 64^3, one MeshBlock, "
                          "VL2+PLM+HLLC, tlim=1", 64 ** 3, 1)
    if a.config == "config1":   # MHD linear wave 128^3 in 32^3 MeshBlocks, tlim=0.5
        b = base("LinWave", "vl2", 2, 0.5, -1)
        b.update(mesh_block((128, 128, 128), (0.0, 0.0, 0.0), (3.0, 1.5, 1.5), (32, 32, 32)))
        b["hydro"] = {"gamma": 1.666666666666667, "iso_sound_speed": 1.0}
        b["problem"] = {"compute_error": "false", "wave_flag": 0, "amp": 1.0e-6, "vflow": 0.0}
        return run_binary(b, ["--prob=linear_wave", "-b"], "BASELINE configs[1]: MHD linear wave 128^3, 32^3 "
                          "MeshBlocks, VL2+PLM+HLLD+CT, tlim=0.5", 128 ** 3, 1)
    if a.config == "config3":   # KH hydro, RK3 + PPM + HLLC + one passive scalar, 64^3 blocks
        n = a.per_gpu
        b = base("kh", "rk3", 3, 10.0, a.cycles)
        b.update(mesh_block((n, n, n), (-0.5, -1.0, -0.25), (0.5, 1.0, 0.25), (a.block,) * 3))
        b["hydro"] = {"gamma": 1.6666667, "iso_sound_speed": 1.0}
        b["problem"] = {"iprob": 4, "amp": 0.01, "vflow": 1.0, "vboost": 0.0, "drho_rho0": 0.0}
        return run_binary(b, ["--prob=kh", "--nghost=3", "--nscalars=1"], "BASELINE configs[3] shape: KH hydro %d^3, "
                          "RK3+PPM+HLLC + 1 passive scalar, %d^3 MeshBlocks" % (n, a.block), n ** 3, 1)
    if a.config == "sr":   # special-relativistic hydro blast (pressure ratio 1e4), VL2 + PLM + HLLC (hllc_rel.cpp)
        n = a.per_gpu
        b = base("Blast", "vl2", 2, 10.0, a.cycles)
        b.update(mesh_block((n, n, n), (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), (a.block,) * 3))
        b["hydro"] = {"gamma": 1.666666666667}
        b["problem"] = {"pamb": 0.01, "prat": 1.0e4, "radius": 0.1, "damb": 1.0, "drat": 1.0, "ramp": 0.0, "b0": 0.0,
                        "angle": 0}
        return run_binary(b, ["--prob=blast", "-s"], "SR hydro blast %d^3, %d^3 MeshBlocks, VL2+PLM+HLLC (-s)" % (n, a.block),
                          n ** 3, 1)
    if a.config == "smr":   # MHD blast, 128^3 root mesh in 32^3 MeshBlocks, the central 2x2x2 root blocks refined once
        b = base("Blast", "vl2", 2, 10.0, a.cycles)
        b.update(mesh_block((128, 128, 128), (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), (32, 32, 32), num_gpus=a.gpus,
                            refinement="static"))
        b["refinement1"] = {"x1min": -0.4, "x1max": 0.4, "x2min": -0.4, "x2max": 0.4, "x3min": -0.4, "x3max": 0.4, "level": 1}
        b["hydro"] = {"gamma": 1.666666666667, "iso_sound_speed": 0.4082482905}
        b["problem"] = {"pamb": 0.1, "prat": 100.0, "radius": 0.1, "damb": 1.0, "drat": 1.0, "ramp": 0.0, "b0": 1.0,
                        "angle": 30}
        nblocks = 64 - 8 + 64
        return run_binary(b, ["--prob=blast", "-b"], "MHD blast on a refined mesh: 128^3 root, 32^3 MeshBlocks, centre "
                          "refined once (%d blocks), VL2+PLM+HLLD+CT" % nblocks, nblocks * 32 ** 3, a.gpus)
    # the bench workload: MHD blast, 256^3 zones per GPU in 64^3 MeshBlocks, weak-scaled
    f = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[a.gpus]
    mesh = tuple(a.per_gpu * x for x in f)
    b = base("Blast", "vl2", 2, 10.0, a.cycles)
    b.update(mesh_block(mesh, tuple(-1.0 * x for x in f), tuple(1.0 * x for x in f), (a.block,) * 3, num_gpus=a.gpus))
    b["hydro"] = {"gamma": 1.666666666667, "iso_sound_speed": 0.4082482905}
    b["problem"] = {"pamb": 0.1, "prat": 100.0, "radius": 0.1, "damb": 1.0, "drat": 1.0, "ramp": 0.0, "b0": 1.0,
                    "angle": 30}
    run_binary(b, ["--prob=blast", "-b"], "MHD blast %dx%dx%d, %d^3 MeshBlocks, VL2+PLM+HLLD+CT" % (mesh + (a.block,)),
               mesh[0] * mesh[1] * mesh[2], a.gpus)


if __name__ == "__main__":
    main()
