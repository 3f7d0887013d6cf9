"""Turns an .ncu-rep (ncu --set full) into the text summary committed under profiles/ and, with
--traffic, into profiles/ncu_traffic.json (per-launch DRAM bytes + FP64 pipe % that bench.py quotes).
usage: python tools/ncu_summary.py report.ncu-rep out.txt [--traffic profiles/ncu_traffic.json]"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_config_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
NAMES = {"k_sweep_tile_staged<1,": "sweep_x2", "k_sweep_tile_staged<2,": "sweep_x3", "k_sweep_x1<": "sweep_x1", "k_sweep_x1_cells<": "sweep_x1",
         "CornerEPoint": "corner_e", "Cons2PrimPoint": "cons2prim", "CTIntegratePoint": "ct_integrate"}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none (B200, 256^3 zones, 64^3 MeshBlocks, MHD PLM stage of the fused VL2 cycle)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write("## %s\n" % name)
            vals = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    vals[k] = r[i]
                    f.write("  %-70s %s %s\n" % (k, r[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
                    try:
                        stalls.append((float(r[i]), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
                    except ValueError:
                        pass
            tot = sum(v for v, _ in stalls) or 1.0
            f.write("  warp-stall samples (share of all samples):\n")
            for v, h in sorted(stalls, reverse=True)[:7]:
                f.write("      %5.1f %%  %s\n" % (100.0 * v / tot, h))
            for pat, key in NAMES.items():
                if pat in name.replace(" ", "").replace("<1,", "<1,").replace("(int)", "") or pat in name:
                    def gb(k):
                        i = hdr.index(k)
                        v = float(r[i])
                        u = units[i].lower()
                        return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)
                    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
                    traffic[key] = {"traffic": rd + wr, "traffic_read": rd, "traffic_write": wr,
                                    "fp64_pipe_pct": float(vals["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]),
                                    "ms_under_ncu": float(vals["gpu__time_duration.sum"]),
                                    "ncu_kernel": name, "source": out + " (ncu --set full, 256^3, PLM stage)"}
    if "--traffic" in sys.argv:
        with open(sys.argv[sys.argv.index("--traffic") + 1], "w") as f:
            json.dump(traffic, f, indent=1)
    print(open(out).read())


if __name__ == "__main__":
    main()
