"""The C++ host binary (Mesh / MeshBlock / TaskList API) on the bench workload: MHD blast, 256^3
zones per GPU in 64^3 MeshBlocks, one process driving N GPUs (<mesh> num_gpus = N).  Prints one
JSON line with the binary's own zone-cycles/gpu_wsecond (main loop wall clock, as the reference
prints zone-cycles/wsecond).  usage: python tools/host_bench.py [--gpus N] [--cycles K]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402  (athinput formatting only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cycles", type=int, default=30)
    ap.add_argument("--per-gpu", type=int, default=256)
    ap.add_argument("--block", type=int, default=64)
    a = ap.parse_args()
    f = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[a.gpus]
    mesh = tuple(a.per_gpu * x for x in f)
    b = refrun.blast_blocks(mesh=mesh, block=(a.block,) * 3, tlim=10.0)
    b["time"]["nlim"] = a.cycles
    b["time"]["ncycle_out"] = 10
    b["output1"]["dt"] = 1.0e30
    b["mesh"]["num_gpus"] = a.gpus
    for d, n in zip("123", f):   # keep dx equal in every direction
        b["mesh"]["x%smin" % d], b["mesh"]["x%smax" % d] = -1.0 * n, 1.0 * n
    del b["output2"]
    wd = tempfile.mkdtemp(prefix="athhostbench_")
    inp = os.path.join(wd, "athinput")
    open(inp, "w").write(refrun.format_athinput(b))
    exe = os.path.join(ROOT, "athena_b200", "bin", "athena_b200")
    r = subprocess.run([exe, "-i", inp, "-d", wd, "--prob=blast", "-b"], capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:], file=sys.stderr)
        sys.exit(1)
    zc = float(re.search(r"zone-cycles/gpu_wsecond = (\S+)", r.stdout).group(1))
    wt = float(re.search(r"gpu wtime used = (\S+)", r.stdout).group(1))
    print(json.dumps({"host": "athena_b200/bin/athena_b200 (C++ Mesh/TaskList host, one process, <mesh> num_gpus=%d)" % a.gpus,
                      "metric": "zone-cycles/gpu_wsecond (main loop wall clock)", "value": zc, "n_gpus": a.gpus,
                      "mesh": list(mesh), "meshblock": a.block, "cycles": a.cycles, "main_loop_s": wt,
                      "ms_per_cycle": 1e3 * wt / a.cycles}))
    subprocess.run(["rm", "-rf", wd])


if __name__ == "__main__":
    main()
