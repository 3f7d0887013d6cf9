"""Formats an `ncu --metrics gpu__time_duration.sum,... --csv` launch list into the per-launch table
committed under profiles/.  usage: python tools/ncu_launch_list.py launches.csv out.txt [traffic.json]
With a third argument it also writes the DRAM bytes of ONE whole VL2 cycle (sum over its launches)
and per kernel, keyed the way bench.py's breakdown names them (bench.py reads that file for
roofline.traffic / roofline.dominant)."""
import csv
import json
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

NAMES = [("k_sweep_march<1, 1,", "sweep_x2_donor"), ("k_sweep_march<2, 1,", "sweep_x3_donor"),
         ("k_sweep_march<1, 2,", "sweep_x2_plm"), ("k_sweep_march<2, 2,", "sweep_x3_plm"),
         ("k_sweep_x1_rows<1,", "sweep_x1_donor_integrate"), ("k_sweep_x1_rows<2,", "sweep_x1_plm_integrate"),
         ("CornerEPoint", "athk_compute_corner_e"), ("CTIntegratePoint", "ct_integrate"),
         ("k_points_min<Cons2PrimPoint", "athk_cons2prim_dt"), ("k_points<Cons2PrimPoint", "athk_cons2prim"),
         ("GhostFill", "athk_exchange_ghosts"), ("EmfPackPoint", "athk_send_flux_correction"),
         ("EmfApplyPoint", "athk_recv_flux_correction")]


def bench_name(kernel):
    k = kernel.replace("athk::", "").replace("(int)", "").replace("(bool)", "")
    for pat, name in NAMES:
        if pat in k:
            return name
    return None


if len(sys.argv) > 3:
    # one cycle = from a stage-1 x2 donor sweep to the launch before the next one
    starts = [n for n, k in enumerate(order) if bench_name(per[k]["name"]) == "sweep_x2_donor"]
    if len(starts) < 3:
        raise SystemExit("launch list holds fewer than two complete cycles")
    a, b = starts[0], starts[1]   # (the first complete cycle of the capture window)
    cyc = [per[k] for k in order[a:b]]
    kern = {}
    for p in cyc:
        nm = bench_name(p["name"]) or p["name"][:40]
        d = kern.setdefault(nm, {"launches_per_cycle": 0, "traffic_read": 0.0, "traffic_write": 0.0, "us_under_ncu": 0.0,
                                 "fp64_pipe_pct": 0.0})
        d["launches_per_cycle"] += 1
        d["traffic_read"] += 1e6 * p.get("dram__bytes_read.sum", 0.0)
        d["traffic_write"] += 1e6 * p.get("dram__bytes_write.sum", 0.0)
        d["us_under_ncu"] += p.get("gpu__time_duration.sum", 0.0)
        d["fp64_pipe_pct"] = max(d["fp64_pipe_pct"], p.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0.0))
    for d in kern.values():   # per LAUNCH, like roofline.achieved
        n = d["launches_per_cycle"]
        d["traffic"] = (d["traffic_read"] + d["traffic_write"]) / n
        d["traffic_read"] /= n
        d["traffic_write"] /= n
        d["us_under_ncu"] /= n
    out = {"cycle_dram_bytes": sum(1e6 * (p.get("dram__bytes_read.sum", 0.0) + p.get("dram__bytes_write.sum", 0.0)) for p in cyc),
           "cycle_us_under_ncu": sum(p.get("gpu__time_duration.sum", 0.0) for p in cyc),
           "launches_per_cycle": len(cyc), "kernels": kern,
           "source": "%s (ncu --clock-control none, launches %d..%d = one VL2 cycle)" % (sys.argv[2], a, b - 1)}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print("cycle: %d launches, %.2f GB DRAM, %.2f ms under ncu" % (len(cyc), out["cycle_dram_bytes"] / 1e9,
                                                                  out["cycle_us_under_ncu"] / 1e3))
