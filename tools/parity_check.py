"""Development tool: run golden fixtures through a given build of the library (fused and
operator paths) and report bit-exactness.  usage: python tools/parity_check.py lib.so [case ...]"""
import ctypes
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from athena_b200 import capi  # noqa: E402

lib = capi.bind(ctypes.CDLL(os.path.abspath(sys.argv[1])))
cases = sys.argv[2:] or sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(util.GOLDEN, "*.npz")))
bad = 0
for name in cases:
    for fused in (True, False):
        try:
            meta, z, ctx, ti = util.run_golden_case(lib, name, fused=fused)
            util.assert_matches_golden(meta, z, ctx, ti)
            ctx.close()
            res = "ok"
        except AssertionError as ex:
            res = "MISMATCH " + str(ex)[:120].replace("\n", " ")
            bad += 1
        print("%-45s fused=%-5s %s" % (name, fused, res), flush=True)
print("parity_check: %d case(s), %d mismatch(es)" % (2 * len(cases), bad))
sys.exit(1 if bad else 0)
