"""Development tool: per-kernel stall summary from an .ncu-rep (source page, SASS).
usage: python tools/ncu_stalls.py report.ncu-rep [ntop]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        kern.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
seen = set()
for k in kern:
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    idx = {n: i for i, n in enumerate(k["hdr"])}
    stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[idx["# Samples"]]) for r in k["rows"])
    print("##", k["name"][:100], "samples", tot, "static instrs", len(k["rows"]))
    agg = {n: sum(int(r[idx[n]]) for r in k["rows"]) for n in stall_cols}
    print("   stalls:", ", ".join("%s=%.1f%%" % (n[6:], 100.0 * v / max(tot, 1))
                                  for n, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    ex = sum(int(r[idx["Instructions Executed"]]) for r in k["rows"])
    fam = {}
    for r in k["rows"]:
        toks = r[1].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        fam[op] = fam.get(op, 0) + int(r[idx["Instructions Executed"]])
    print("   executed mix:", ", ".join("%s=%.1f%%" % (o, 100.0 * v / ex)
                                        for o, v in sorted(fam.items(), key=lambda x: -x[1])[:14]))
    for col in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier"):
        top = sorted(k["rows"], key=lambda r: -int(r[idx[col]]))[:ntop]
        print("   top", col)
        for r in top:
            if int(r[idx[col]]) == 0:
                break
            print("      %6s %6s  %s" % (r[idx[col]], r[idx["# Samples"]], r[1].strip()[:80]))
