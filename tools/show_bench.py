"""Development helper: one-line summary of a bench.py JSON line. usage: show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4e %s  ms/step %.3f  e2e %.3e  launches %s" % (d["value"], d["unit"], d["ms_per_step"],
                                                            d["e2e"]["value"], d.get("gpu_launches")))
bd = d.get("roofline", {}).get("breakdown_ms_per_cycle", {})
print("  ".join("%s=%.3f" % kv for kv in bd.items()))
