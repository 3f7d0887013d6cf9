"""Formats an `ncu --metrics gpu__time_duration.sum,... --csv` launch list into the per-launch table
committed under profiles/.  usage: python tools/ncu_launch_list.py launches.csv out.txt"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 or (r and r[0] == "ID")]
hdr = rows[0]
ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
per = {}
order = []
for r in rows[1:]:
    k = r[ix["ID"]]
    if k not in per:
        per[k] = {"name": r[ix["Kernel Name"]]}
        order.append(k)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]].lower()
    if u in ("ns", "nsecond"):
        v /= 1e3
    elif u in ("ms", "msecond"):
        v *= 1e3
    elif u.startswith("gbyte"):
        v *= 1e3
    elif u.startswith("kbyte"):
        v /= 1e3
    elif u == "byte":
        v /= 1e6
    per[k][r[ix["Metric Name"]]] = v
with open(sys.argv[2], "w") as f:
    f.write("# one fused VL2 cycle and a half, 256^3 zones in 64^3 MeshBlocks on one B200; ncu --clock-control none\n")
    f.write("# (per-launch times under ncu are serialised and cold-cache: compare SHARES with bench.py's breakdown)\n")
    for n, k in enumerate(order):
        p = per[k]
        f.write("%3d %-46s %8.1f us fp64=%5.1f%% issue=%5.1f%% rd=%7.1fMB wr=%7.1fMB warps=%4.1f%%\n" % (
            n, p["name"].replace("void athk::", "").replace("athk::", "")[:46], p.get("gpu__time_duration.sum", 0.0),
            p.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0.0),
            p.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0),
            p.get("dram__bytes_read.sum", 0.0), p.get("dram__bytes_write.sum", 0.0),
            p.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)))
print(open(sys.argv[2]).read()[:3500])
