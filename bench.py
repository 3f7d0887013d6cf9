#!/usr/bin/env python
"""bench.py -- zone-cycles/sec of the 3-D MHD VL2+PLM+HLLD+CT MeshBlock update on B200.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1 under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

Workload: BASELINE.json configs[2] weak-scaled -- MHD blast wave, 256^3 zones per GPU in 64^3
MeshBlocks (64 blocks per GPU), periodic, gamma=5/3, cfl 0.3; a "step" is one full VL2 cycle
(2 stages: donor-cell predictor + PLM corrector, HLLD, corner-E, EMF correction, CT, ghost
exchange, W(U), new dt) over every MeshBlock.  Metric as the reference prints it
(/root/reference/src/main.cpp:536,639-651): zone-cycles / second.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_MHD_VL2 = 320.0   # algorithmic bytes per zone-cycle, SURVEY.md section 8(d)
F_ALG_MHD_VL2 = 3600.0  # algorithmic FP64 flops per zone-cycle (div/sqrt as 8-10), SURVEY.md section 8(d)
FP64_PEAK_TFLOPS = 36.0  # fallback: 2 flops x 18.0 T FP64 FMA thread-instructions/s (round-1 measurement)


def fp64_peak():
    """FP64 FMA peak of this pool's B200s from the committed raw output of tools/micro/fp64_peak.cu
    (profiles/r2_fp64_peak.json: best rate over resident warps x ILP, with the clock record)."""
    p = os.path.join(ROOT, "profiles", "r2_fp64_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return 2.0 * float(d["best_t_thread_inst_per_s"]), "profiles/r2_fp64_peak.json (tools/micro/fp64_peak.cu)"
    return FP64_PEAK_TFLOPS, "fallback: 18.0 T DFMA thread-instr/s (round 1)"


def committed_traffic():
    """Whole-cycle and per-kernel DRAM bytes from the committed ncu launch list of this bench
    (profiles/r5_cycle_traffic.json -- the library as it stands at the end of round 2; r2_cycle_traffic.json is
    the same capture from the middle of the round -- made by tools/ncu_launch_list.py)."""
    for name in ("r5_cycle_traffic.json", "r2_cycle_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p))
    return None
# algorithmic bytes per zone of ONE launch of the named kernel (kernel table)
DOMINANT_ALG_BYTES = {"sweep_x1": 31 * 8, "sweep_x2": 16 * 8, "sweep_x3": 16 * 8}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled through
    NVML every few ms (nvidia-smi itself takes ~0.5 s per query, longer than a short timed
    region); falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev = dev
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self.stop_flag = False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[dev]) if vis and vis.split(",")[dev].isdigit() else dev
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        if self.h is None or not self.sm:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=clocks.sm,clocks.max.sm",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
                a, b = [float(x) for x in r.stdout.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi (idle)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm), "power_w_max": max(self.power) if self.power else None,
                "source": "nvml"}


def weak_mesh(ngpu, per_gpu):
    """256^3 per GPU: 1 -> 256^3, 2 -> 512x256x256, 4 -> 512x512x256, 8 -> 512^3."""
    f = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[ngpu]
    return tuple(per_gpu * x for x in f), f


def zorder_partition(nbtotal, nranks):
    """Contiguous Z-order ranges per rank (mesh/amr_loadbalance.cpp:73-131, equal costs)."""
    base, rem = divmod(nbtotal, nranks)
    out, s = [], 0
    for r in range(nranks):
        n = base + (1 if r >= nranks - rem else 0)   # remainder assigned from the end
        out.append(list(range(s, s + n)))
        s += n
    return out


def run_reference(args):
    """Reference arm: the UNMODIFIED reference binary (oracle/_ref) on the host cores, a bounded
    sample of the same workload (same physics/config; smaller mesh so it ends in minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refrun
    cfg = "blast_mhd_hlld"
    cores = os.cpu_count() or 1
    n = args.ref_mesh
    blk = args.ref_block
    nblocks = (n // blk) ** 3
    threads = max(1, min(cores, nblocks))
    blocks = blast_athinput(n, blk, nlim=1)
    vals = []
    t0 = time.time()
    nrun = args.warmup + args.steps
    # each "step" of the reference arm is a bounded sample: ref_cycles cycles of the n^3 mesh
    for it in range(nrun):
        blocks["time"]["nlim"] = args.ref_cycles
        wd, out = refrun.run_reference(cfg, blocks, threads=threads)
        zc = refrun.parse_zone_cycles(out)
        subprocess.run(["rm", "-rf", wd])
        if it >= args.warmup:
            vals.append(zc["zone-cycles/omp_wsecond"])
        if time.time() - t0 > 240 and vals:
            break
    v = sorted(vals)[len(vals) // 2]
    sample = "%d^3 mesh, %d^3 MeshBlocks, %d cycles per step, %d OpenMP threads" % (
        n, blk, args.ref_cycles, threads)
    line = {"impl": "reference", "metric": "zone-cycles/sec (3D MHD VL2+PLM+HLLD+CT)", "value": v,
            "unit": "zone-cycles/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
            "ms_per_step": 1e3 * (n ** 3) * args.ref_cycles / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args), ref_mesh=n, ref_block=blk, ref_cycles_per_step=args.ref_cycles,
                           note="same problem, physics and algorithm as the product arm; the CPU arm times a "
                                "bounded sample on a %d^3 mesh in %d^3 MeshBlocks (zone-cycles/s is size-"
                                "independent to first order, SURVEY 8d)" % (n, blk)),
            "cpu_baseline": {"value": v, "unit": "zone-cycles/s", "cores": threads,
                             "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "zone-cycles/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit(line)


def blast_athinput(n, blk, nlim):
    return {
        "job": {"problem_id": "Blast"},
        "output1": {"file_type": "hst", "dt": 1.0e30},
        "time": {"cfl_number": 0.3, "nlim": nlim, "tlim": 1.0, "integrator": "vl2", "xorder": 2,
                 "ncycle_out": 1000},
        "mesh": {"nx1": n, "x1min": -1.0, "x1max": 1.0, "ix1_bc": "periodic", "ox1_bc": "periodic",
                 "nx2": n, "x2min": -1.0, "x2max": 1.0, "ix2_bc": "periodic", "ox2_bc": "periodic",
                 "nx3": n, "x3min": -1.0, "x3max": 1.0, "ix3_bc": "periodic", "ox3_bc": "periodic",
                 "num_threads": 1},
        "meshblock": {"nx1": blk, "nx2": blk, "nx3": blk},
        "hydro": {"gamma": 1.666666666667, "iso_sound_speed": 0.4082482905},
        "problem": {"pamb": 0.1, "prat": 100.0, "radius": 0.1, "damb": 1.0, "drat": 1.0,
                    "ramp": 0.0, "b0": 1.0, "angle": 30},
    }


def workload_config(args):
    return {"workload": "mhd_blast %d^3 zones per GPU, %d^3 MeshBlocks, VL2+PLM+HLLD+CT, periodic "
                        "(BASELINE configs[2], weak-scaled)" % (args.per_gpu, args.block),
            "zones_per_gpu": args.per_gpu ** 3, "meshblock": args.block, "integrator": "vl2",
            "xorder": 2, "rsolver": "hlld", "nghost": 2, "gamma": 1.666666666667,
            "cfl_number": 0.3, "l2_policy": "inputs larger than L2 (state >= 1 GB per GPU)",
            "parallelism": "meshblock-z-order x%d" % args.gpus}


def cpu_baseline_sample(args):
    """Rank 0, N=1: the reference binary on the box's host cores for ~10-30 s of CPU work."""
    from oracle import refrun
    cfg = "blast_mhd_hlld"
    if not refrun.have_ref(cfg):
        return None
    cores = os.cpu_count() or 1
    n, blk = args.ref_mesh, args.ref_block
    threads = max(1, min(cores, (n // blk) ** 3))
    # aim for ~15 s: the reference does ~0.6e6 zone-cycles/s/core
    cyc = max(2, min(40, int(15.0 * 0.6e6 * threads / n ** 3)))
    blocks = blast_athinput(n, blk, nlim=cyc)
    wd, out = refrun.run_reference(cfg, blocks, threads=threads)
    subprocess.run(["rm", "-rf", wd])
    zc = refrun.parse_zone_cycles(out)
    return {"value": zc["zone-cycles/omp_wsecond"], "unit": "zone-cycles/s", "cores": threads,
            "kind": "reference",
            "sample": "%d^3 mesh, %d^3 MeshBlocks, %d cycles, %d OpenMP threads, g++ -O3" % (
                n, blk, cyc, threads)}


def run_ours(args):
    import ctypes as C

    import numpy as np
    import torch

    from athena_b200 import capi, problems, pyhost

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = capi.load()
    parity = None
    if world > 1 and not args.no_parity_check:
        # before anything is timed: a golden fixture (the reference's own output) split over the
        # ranks -- the multi-GPU exchange must reproduce it bit for bit
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import multi_check
        parity = multi_check.split_golden_case(lib, args.parity_case, rank, world, local_rank)
    mesh_nx, _ = weak_mesh(world, args.per_gpu)
    mesh = pyhost.UniformMesh(mesh_nx, (args.block,) * 3, (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), ng=2)
    parts = zorder_partition(mesh.nbtotal, world)
    rank_of = {}
    for r, gl in enumerate(parts):
        for g in gl:
            rank_of[g] = r
    gids = parts[rank]
    gamma = 1.666666666667
    ctx = pyhost.Context(mesh, True, gamma, device=local_rank, gids=gids, rank_of=rank_of,
                         my_rank=rank, lib=lib)
    stream = torch.cuda.ExternalStream(lib.athk_stream(ctx.h))
    if world > 1:
        from athena_b200 import multigpu
        multigpu.attach(ctx, rank, world, device="cuda", stream=stream)

    # ---- synthetic ICs on the host (pinned), reference pgen formulas --------------------------
    nb = len(gids)
    shp = {k: ctx.shape(k) for k in ("U", "B_X1F", "B_X2F", "B_X3F")}
    host = {k: torch.empty((nb,) + shp[k], dtype=torch.float64, pin_memory=True) for k in shp}
    hnp = {k: v.numpy() for k, v in host.items()}
    for lb, g in enumerate(gids):
        ic = problems.blast_block(mesh, mesh.loc[g], True, gamma)
        hnp["U"][lb] = ic["u"]
        hnp["B_X1F"][lb] = ic["b1"]
        hnp["B_X2F"][lb] = ic["b2"]
        hnp["B_X3F"][lb] = ic["b3"]
    dp = C.POINTER(C.c_double)
    for k in hnp:   # initial state: AthenaArray extents (ghost zones zero), one transfer per array
        ctx.call("athk_upload_all", capi.ARRAYS[k], hnp[k].ctypes.data_as(dp), 0)
    ctx.sync()
    # e2e buffers: the ACTIVE zone of every MeshBlock (ghost zones are refreshed on the device),
    # pinned; one PCIe transfer per array and direction
    nx = args.block
    ashp = {"U": (5, nx, nx, nx), "B_X1F": (nx, nx, nx + 1), "B_X2F": (nx, nx + 1, nx), "B_X3F": (nx + 1, nx, nx)}
    ahost = {k: torch.empty((nb,) + ashp[k], dtype=torch.float64, pin_memory=True) for k in ashp}
    anp = {k: v.numpy() for k, v in ahost.items()}
    state_bytes = sum(v.numel() * 8 for v in ahost.values())

    def h2d():
        for k in anp:
            ctx.call("athk_upload_all", capi.ARRAYS[k], anp[k].ctypes.data_as(dp), 1)

    def d2h():
        for k in anp:
            ctx.call("athk_download_all_async", capi.ARRAYS[k], anp[k].ctypes.data_as(dp), 1)
        ctx.sync()

    reduce_min = multigpu.reduce_min_fn("cuda") if world > 1 else None
    ti = pyhost.TimeIntegrator(ctx, "vl2", 2, 0.3, fused=not args.unfused)
    ti.initialize(reduce_min=reduce_min)

    zones = args.per_gpu ** 3 * world

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(nsteps):
            fn()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    step = lambda: ti.step(reduce_min=reduce_min)   # noqa: E731
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launches()
    ms = timed(step, args.steps)
    launches = ctx.launches() - l0
    sampler.stop_flag = True
    value = zones * args.steps / (ms * 1e-3)

    # ---- per-kernel breakdown, measured live with CUDA events on the launching stream --------
    # (athk_profile brackets every kernel launch inside the library; the NCCL exchanges are
    # bracketed here).  2 extra cycles after the timed region, so `value` is not perturbed.
    breakdown = {}
    nprof = 2
    evs = []
    xch = getattr(ctx, "xch", None)
    if rank == 0:
        ctx.call("athk_profile", 1, None, 0)
        if xch is not None:
            def wrap(fn, key):
                def g():
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(stream)
                    fn()
                    e.record(stream)
                    evs.append((key, s, e))
                return g
            orig_xg, orig_xe = xch.exchange_ghosts, xch.exchange_emf
            xch.exchange_ghosts = wrap(orig_xg, "nccl_exchange_ghosts")
            xch.exchange_emf = wrap(orig_xe, "nccl_exchange_emf")
    for _ in range(nprof):
        step()
    if rank == 0:
        buf = C.create_string_buffer(16384)
        ctx.call("athk_profile", 0, buf, len(buf))
        rep = json.loads(buf.value.decode() or "{}")
        for k, v in rep.items():
            breakdown[k] = {"ms_per_cycle": v["ms"] / nprof, "launches_per_cycle": v["launches"] / nprof}
        if xch is not None:
            xch.exchange_ghosts, xch.exchange_emf = orig_xg, orig_xe
            torch.cuda.synchronize()
            for name, s_, e_ in evs:
                d = breakdown.setdefault(name, {"ms_per_cycle": 0.0, "launches_per_cycle": 0.0})
                d["ms_per_cycle"] += s_.elapsed_time(e_) / nprof
                d["launches_per_cycle"] += 1.0 / nprof
    barrier()

    # ---- e2e: host buffers in, host buffers out, every step ---------------------------------
    def e2e_step():
        h2d()
        # the state came from the host: redo the start-of-cycle exchange + W(U) the way
        # Mesh::Initialize does after a restart, then one full cycle, then read everything back
        ti.refresh_ghosts_and_primitives()
        ti.step(reduce_min=reduce_min)
        d2h()
    d2h()   # make the host copy the current state so that repeated e2e steps stay physical
    e2e_step()
    e2e_n = max(2, min(args.steps, 5))
    ms_e2e = timed(e2e_step, e2e_n)
    e2e_value = zones * e2e_n / (ms_e2e * 1e-3)

    # ---- e2e with several MeshBlock sets in flight on the GPU (N=1): further, independent contexts
    # of the same workload (own stream, own pinned buffers, own host thread -- the C ABI is per
    # context).  Every step of every set still is host -> device, exchange + W(U), one cycle,
    # device -> host; the H2D copy of one set runs under the cycle of another and the D2H copy of a
    # third (PCIe is full duplex, the copy engines and the SMs are separate units), which is how a
    # host that advances several meshes (an ensemble, a parameter scan) would drive the library.
    e2e_pipe = None
    nsets = max(1, args.e2e_sets)
    if world == 1 and nsets > 1 and args.per_gpu <= 320:
        import time as _time

        lanes = [(ctx, ti, anp, stream, ahost)]
        for _ in range(nsets - 1):
            cx = pyhost.Context(mesh, True, gamma, device=local_rank, gids=gids, rank_of=rank_of,
                                my_rank=rank, lib=lib)
            for k in hnp:
                cx.call("athk_upload_all", capi.ARRAYS[k], hnp[k].ctypes.data_as(dp), 0)
            tx = pyhost.TimeIntegrator(cx, "vl2", 2, 0.3, fused=not args.unfused)
            tx.initialize()
            xhost = {k: torch.empty((nb,) + ashp[k], dtype=torch.float64, pin_memory=True) for k in ashp}
            xnp = {k: v.numpy() for k, v in xhost.items()}
            for k in xnp:   # every set starts from its own current state
                cx.call("athk_download_all_async", capi.ARRAYS[k], xnp[k].ctypes.data_as(dp), 1)
            cx.sync()
            lanes.append((cx, tx, xnp, torch.cuda.ExternalStream(lib.athk_stream(cx.h)), xhost))

        # set n starts n/nsets of a step behind the first, so that its H2D copy meets the cycle and
        # the D2H copy of the others and not their H2D copies (sets in lockstep would share each
        # direction of the link and leave the other idle)
        lag_s = ms_e2e / e2e_n * 1e-3 / nsets

        from athena_b200 import ensemble
        ens = ensemble.HostSteppedEnsemble([ensemble.Member(ln[0], ln[1], ln[2], local_rank) for ln in lanes])

        def run_lanes(n_each):
            ens.run(n_each, lag_s)

        def sync_lanes():
            for ln in lanes:
                ln[0].sync()
            torch.cuda.synchronize()

        run_lanes(1)   # warm-up (staging buffers, pinned pages touched)
        sync_lanes()
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in lanes]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in lanes]
        for e, ln in zip(ev0, lanes):
            e.record(ln[3])
        t0 = _time.perf_counter()
        pipe_n = 2 * e2e_n   # steps per set (the stagger is inside the timed region)
        run_lanes(pipe_n)
        for e, ln in zip(ev1, lanes):
            e.record(ln[3])
        sync_lanes()
        wall_ms = (_time.perf_counter() - t0) * 1e3
        dev_ms = max(a.elapsed_time(b) for a in ev0 for b in ev1)
        ms_pipe = max(wall_ms, dev_ms)   # the slower clock of the two
        # the added sets ran the same cycles from the same state: their host copies must agree bit for bit
        same = all(bool(np.array_equal(lanes[1][2][k], ln[2][k])) for ln in lanes[2:] for k in anp) \
            if nsets > 2 else None
        e2e_pipe = {"value": zones * nsets * pipe_n / (ms_pipe * 1e-3), "unit": "zone-cycles/s",
                    "ms_per_step": ms_pipe / (nsets * pipe_n), "steps": nsets * pipe_n, "sets_in_flight": nsets,
                    "device_ms": dev_ms, "wall_ms": wall_ms, "added_sets_bit_identical": same}
        for ln in lanes[1:]:
            ln[0].close()
        del lanes[1:]

    # ---- the fast-arithmetic kernel set on the same workload (reported next to `value`, which is
    # the strict, bit-exact set): same state, a few cycles, device-timed -------------------------
    fast = None
    if world == 1 and not args.no_fast_arith:
        ctx.close()
        ctx = pyhost.Context(mesh, True, gamma, device=local_rank, gids=gids, rank_of=rank_of,
                             my_rank=rank, lib=lib, arith=capi.ARITH_FAST)
        stream = torch.cuda.ExternalStream(lib.athk_stream(ctx.h))
        for k in hnp:
            ctx.call("athk_upload_all", capi.ARRAYS[k], hnp[k].ctypes.data_as(dp), 0)
        tif = pyhost.TimeIntegrator(ctx, "vl2", 2, 0.3, fused=True)
        tif.initialize()
        for _ in range(args.warmup):
            tif.step()
        nfast = max(3, min(args.steps, 10))
        msf = timed(lambda: tif.step(), nfast)
        fast = {"value": zones * nfast / (msf * 1e-3), "unit": "zone-cycles/s", "ms_per_step": msf / nfast,
                "what": "athk_config.arith = ATHK_ARITH_FAST (FMA contraction, ~1 ulp divisions): agrees with "
                        "the reference to rounding (tests/test_fast_arith.py), not bit for bit; `value` is the "
                        "strict set"}

    # ---- the other single-GPU BASELINE shapes through the same code path (device-timed, a few
    # cycles each): configs[0] hydro 64^3 in ONE MeshBlock (VL2+PLM+HLLC), configs[1] MHD 128^3 in 32^3
    # MeshBlocks (VL2+PLM+HLLD+CT).  Parity of both at full size is in tests/ (C++ host vs reference).
    other = None
    if world == 1 and not args.no_other_workloads:
        other = {}
        ctx.close()
        for key, n, blk, mhd_o in (("configs[0]: hydro wave 64^3, one 64^3 MeshBlock, VL2+PLM+HLLC", 64, 64, False),
                                   ("configs[1]: MHD wave 128^3, 32^3 MeshBlocks, VL2+PLM+HLLD+CT", 128, 32, True)):
            m_o = pyhost.UniformMesh((n,) * 3, (blk,) * 3, (0.0, 0.0, 0.0), (3.0, 1.5, 1.5), ng=2)
            c_o = pyhost.Context(m_o, mhd_o, gamma, device=local_rank, lib=lib)
            for lb, g_ in enumerate(c_o.gids):
                ic = problems.sound_wave_block(m_o, m_o.loc[g_], mhd_o, gamma)
                c_o.upload("U", lb, ic["u"])
                if mhd_o:
                    c_o.upload("B_X1F", lb, ic["b1"])
                    c_o.upload("B_X2F", lb, ic["b2"])
                    c_o.upload("B_X3F", lb, ic["b3"])
            stream = torch.cuda.ExternalStream(lib.athk_stream(c_o.h))
            t_o = pyhost.TimeIntegrator(c_o, "vl2", 2, 0.3, fused=True)
            t_o.initialize()
            for _ in range(5):
                t_o.step()
            n_o = 40
            ctx = c_o   # (timed() synchronises through ctx)
            ms_o = timed(lambda: t_o.step(), n_o)
            other[key] = {"value": n ** 3 * n_o / (ms_o * 1e-3), "unit": "zone-cycles/s", "ms_per_cycle": ms_o / n_o,
                          "cycles": n_o, "data": "synthetic plane wave (athena_b200/problems.py)"}
            c_o.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        per_gpu_value = value / world
        achieved = per_gpu_value * B_ALG_MHD_VL2 / 1e9
        tot = sum(v["ms_per_cycle"] for v in breakdown.values()) or 1.0
        top = max(breakdown, key=lambda k: breakdown[k]["ms_per_cycle"]) if breakdown else None
        dominant = None
        if top:
            t = breakdown[top]
            ms_launch = t["ms_per_cycle"] / max(t["launches_per_cycle"], 1e-9)
            # algorithmic bytes of one launch of the final x1 sweep, per zone: reads w(5) + bcc(2)
            # + b.x1f(1) + d2(5) + d3(5) + u(5), writes u(5) + 2 face EMFs + CT weight = 31 doubles
            alg = DOMINANT_ALG_BYTES.get(top.split("_plm")[0].split("_donor")[0], None)
            dominant = {"kernel": top, "ms_per_launch": ms_launch,
                        "share_of_step": t["ms_per_cycle"] / tot,
                        "alg_bytes_per_zone": alg}
            if alg:
                gbs = alg * args.per_gpu ** 3 / (ms_launch * 1e-3) / 1e9
                dominant.update({"achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak})
            ct = committed_traffic()
            if ct:
                for kname, kv in ct.get("kernels", {}).items():
                    if top.startswith(kname):
                        dominant.update(kv)
        ct = committed_traffic()
        fpeak, fpeak_src = fp64_peak()
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                # DRAM bytes of ONE whole VL2 cycle (sum over every launch of the committed ncu list)
                "traffic": ct["cycle_dram_bytes"] if ct else None,
                "traffic_source": ct.get("source") if ct else None,
                "kernel": "whole VL2 cycle = every kernel of the step; %g algorithmic B/zone-cycle "
                          "(SURVEY 8d). The sweeps are FP64-issue bound, not HBM bound: see "
                          "fp64_pipe_pct in profiles/ and DESIGN.md" % B_ALG_MHD_VL2,
                "peak_source": peak_src, "dominant": dominant,
                "breakdown_ms_per_cycle": {k: round(v["ms_per_cycle"], 4) for k, v in sorted(
                    breakdown.items(), key=lambda kv: -kv[1]["ms_per_cycle"])}}
        # the second ceiling SURVEY 8(d) asks for: the fused stage is FP64-pipe bound, not HBM bound
        tflops = per_gpu_value * F_ALG_MHD_VL2 / 1e12
        roof["fp64"] = {"alg_flops_per_zone_cycle": F_ALG_MHD_VL2, "achieved": tflops, "peak": fpeak,
                        "unit": "TFLOP/s", "frac": tflops / fpeak, "peak_source": fpeak_src}
        if fast is not None:
            roof["fast_arith"] = fast
        if other is not None:
            roof["other_workloads"] = other
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_baseline_sample(args)
            except Exception as ex:   # the bench line must still be printed
                cpu = {"error": str(ex)[:200]}
        what = ("per step: pinned-host active zones of (u,b) of every MeshBlock -> device (one transfer "
                "per array), ghost exchange + W(U), one VL2 cycle, (u,b) + block dt -> host")
        e2e_line = {"value": e2e_value, "unit": "zone-cycles/s", "h2d_bytes_per_step": state_bytes * world,
                    "d2h_bytes_per_step": state_bytes * world + 8 * mesh.nbtotal, "what": what + "; one set of "
                    "MeshBlocks, every step waits for its own copies (H2D, cycle, D2H strictly in sequence)"}
        if e2e_pipe is not None:
            # the end-to-end number of the library driven the way its per-context ABI allows: several
            # independent sets of MeshBlocks per GPU, the copies of one under the work of the others
            e2e_line = {"value": e2e_pipe["value"], "unit": "zone-cycles/s",
                        "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes + 8 * mesh.nbtotal,
                        "what": what + "; %d independent sets of MeshBlocks (contexts) in flight per GPU, each "
                        "driven by its own host thread and stream: the copies of one set overlap the cycle and the "
                        "opposite-direction copy of the others; a step = one set's H2D + cycle + D2H; timed by the "
                        "slower of wall clock and CUDA events" % e2e_pipe["sets_in_flight"],
                        "pipeline": e2e_pipe,
                        "one_set_in_sequence": {"value": e2e_value, "unit": "zone-cycles/s",
                                                "ms_per_step": ms_e2e / e2e_n}}
        line = {"metric": "zone-cycles/sec (3D MHD VL2+PLM+HLLD+CT)", "value": value,
                "unit": "zone-cycles/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args), "clocks": sampler.summary(),
                "e2e": e2e_line,
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "final_dt": ti.dt, "cycles_run": ti.ncycle}
        if parity is not None:
            line["parity_check"] = parity
        if xch is not None:
            line["exchange"] = {"transport": "p2p (peer stores over NVLink + flags)" if xch.p2p else "nccl send/recv",
                                "ghost_bytes_per_stage": int(xch.ghost_bytes), "emf_bytes_per_stage": int(xch.emf_bytes),
                                "p2p_timeouts": int(lib.athk_p2p_timeouts(ctx.h))}
        emit(line)
    if xch is not None:
        xch.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL version banner, torch warnings) may write to fd 1; the contract is ONE JSON
    line on stdout, so fd 1 is pointed at stderr and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--per-gpu", type=int, default=256, help="zones per side per GPU")
    ap.add_argument("--block", type=int, default=64, help="MeshBlock size")
    ap.add_argument("--unfused", action="store_true", help="operator-per-kernel path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--e2e-sets", type=int, default=3,
                    help="independent sets of MeshBlocks (contexts) in flight for the e2e number at N=1; 1 = one set, "
                         "copies and cycle strictly in sequence")
    ap.add_argument("--no-fast-arith", action="store_true", help="skip the extra fast-arithmetic measurement (N=1)")
    ap.add_argument("--no-other-workloads", action="store_true",
                    help="skip the BASELINE configs[0] / configs[1] shape measurements (N=1)")
    ap.add_argument("--parity-case", default="mhd_linwave_vl2_plm_hlld",
                    help="golden fixture run split over the ranks before the timed region (N > 1)")
    ap.add_argument("--ref-mesh", type=int, default=128)
    ap.add_argument("--ref-block", type=int, default=32)
    ap.add_argument("--ref-cycles", type=int, default=4)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
