"""debug driver: python gpurun_tmp/smr_dbg.py <case-name> -- compares host (emul) against the reference, verbosely"""
import glob, os, sys, shutil
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import refrun
import test_host_binary as T

def cases():
    R = T.refined
    return {
      "lw_centre": ("linwave_mhd_hlld", "linear_wave", True, lambda: R(
          refrun.linwave_blocks(True, mesh=(32, 16, 16), block=(8, 8, 8), tlim=0.05, rst_dt=0.025),
          (0.75, 2.25, 0.375, 1.125, 0.375, 1.125, 1))),
      "lw_init": ("linwave_mhd_hlld", "linear_wave", True, lambda: R(
          refrun.linwave_blocks(True, mesh=(32, 16, 16), block=(8, 8, 8), tlim=0.0, rst_dt=0.025),
          (0.75, 2.25, 0.375, 1.125, 0.375, 1.125, 1))),
      "lw_2lev": ("linwave_mhd_hlld", "linear_wave", True, lambda: R(
          refrun.linwave_blocks(True, mesh=(24, 16, 16), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, integrator="rk3"),
          (0.0, 0.6, 0.0, 0.5, 0.9, 1.5, 2))),
      "walls": ("blast_mhd_hlld", "blast", True, lambda: R(
          refrun.blast_blocks(mesh=(32, 32, 32), block=(8, 8, 8), tlim=0.04, rst_dt=0.02, radius=0.25,
                              bcs=["reflecting", "outflow", "outflow", "reflecting", "reflecting", "outflow"],
                              center=(-0.6, 0.5, -0.55)),
          (-1.0, -0.4, 0.3, 1.0, -1.0, -0.3, 1), (0.5, 1.0, -1.0, -0.6, 0.4, 1.0, 2))),
      "blast": ("blast_mhd_hlld", "blast", True, lambda: R(
          refrun.blast_blocks(mesh=(32, 32, 32), block=(8, 8, 8), tlim=0.06, rst_dt=0.03, radius=0.25),
          (-0.3, 0.3, -0.3, 0.3, -0.3, 0.3, 1))),
    }

def main():
    name = sys.argv[1]
    cfg, prob, mhd, mk = cases()[name]
    blocks = mk()
    if len(sys.argv) > 2:
        blocks["time"]["nlim"] = int(sys.argv[2])
    wd_ref, _ = refrun.run_reference(cfg, blocks)
    hb = {k: dict(v) for k, v in blocks.items()}
    if os.environ.get("SMR_NGPU"):
        hb["mesh"]["num_gpus"] = int(os.environ["SMR_NGPU"])
    wd, r = T.run_host(T.CUDA_BIN if os.environ.get("SMR_GPU") else T.EMUL_BIN, hb, prob, mhd)
    print("host rc", r.returncode)
    if r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-3000:]); return
    ref_files = sorted(os.path.basename(f) for f in glob.glob(wd_ref + "/*.rst"))
    our_files = sorted(os.path.basename(f) for f in glob.glob(wd + "/*.rst"))
    print(ref_files, our_files)
    bad = 0
    for fn in ref_files:
        if fn not in our_files: continue
        a, c = refrun.read_rst(os.path.join(wd_ref, fn)), refrun.read_rst(os.path.join(wd, fn))
        print(fn, "ncycle", a["ncycle"], c["ncycle"], "time", a["time"], c["time"], "dt", a["dt"], c["dt"], "nblocks", len(a["u"]))
        for k in ("u", "b1", "b2", "b3"):
            nbad = 0
            for blk, (x, y) in enumerate(zip(a[k], c[k])):
                if not np.array_equal(x, y):
                    nbad += 1
                    if nbad <= 3:
                        d = np.argwhere(x != y)
                        print("   ", k, "block", blk, "loc", a["loc"][blk], "nmismatch", len(d), "first", d[:4].tolist(), "last", d[-2:].tolist(),
                              "maxabs", np.abs(x - y).max())
            print("  ", k, "blocks differing:", nbad)
            bad += nbad
    print("TOTAL bad", bad)
    shutil.rmtree(wd, ignore_errors=True); shutil.rmtree(wd_ref, ignore_errors=True)
main()
