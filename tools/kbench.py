"""Operator micro-benchmark on the GPU (development tool): times individual C-ABI calls with CUDA
events on the context stream.  usage: python tools/kbench.py [lib.so ...] [--n 256] [--block 64]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from athena_b200 import capi, problems, pyhost  # noqa: E402


def setup(lib, n, blk, mhd=True, config="blast"):
    kh = config == "kh"   # BASELINE configs[3] shape: hydro, RK3 + PPM + HLLC + one passive scalar
    if kh:
        mhd = False
    mesh = pyhost.UniformMesh((n,) * 3, (blk,) * 3, (-1.0,) * 3, (1.0,) * 3, ng=3 if kh else 2)
    ctx = pyhost.Context(mesh, mhd, 1.666666666667, lib=lib, nscalars=1 if kh else 0)
    for lb, g in enumerate(ctx.gids):
        ic = problems.blast_block(mesh, mesh.loc[g], mhd, 1.666666666667)
        ctx.upload("U", lb, ic["u"])
        if kh:
            ctx.upload("S", lb, 0.5 * ic["u"][0:1])
        if mhd:
            ctx.upload("B_X1F", lb, ic["b1"])
            ctx.upload("B_X2F", lb, ic["b2"])
            ctx.upload("B_X3F", lb, ic["b3"])
    ti = pyhost.TimeIntegrator(ctx, "rk3" if kh else "vl2", 3 if kh else 2, 0.3, fused=True)
    ti.initialize()
    for _ in range(3):
        ti.step()
    return ctx, ti


def time_call(ctx, stream, fn, reps=5):
    fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--block", type=int, default=64)
    ap.add_argument("--sweeps", action="store_true", help="only the fused cycle + its per-kernel breakdown")
    ap.add_argument("--config", default="blast", choices=["blast", "kh"])
    ap.add_argument("--all", action="store_true", help="with --sweeps: list every kernel, not only the sweeps")
    args = ap.parse_args()
    libs = args.libs or [capi.LIB_PATH]
    torch.cuda.init()
    for path in libs:
        lib = capi.bind(ctypes.CDLL(os.path.abspath(path)))
        ctx, ti = setup(lib, args.n, args.block, config=args.config)
        stream = torch.cuda.ExternalStream(lib.athk_stream(ctx.h))
        dt = ti.dt
        res = {}
        if args.sweeps:
            import json
            res["cycle_fused"] = time_call(ctx, stream, lambda: ti.step(), reps=3)
            ctx.call("athk_profile", 1, None, 0)
            for _ in range(2):
                ti.step()
            buf = ctypes.create_string_buffer(16384)
            ctx.call("athk_profile", 0, buf, len(buf))
            rep = json.loads(buf.value.decode() or "{}")
            for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
                if "sweep" in k or args.all:
                    res[k] = v["ms"] / 2
            print(os.path.basename(path), " ".join("%s=%.3f" % kv for kv in res.items()), flush=True)
            ctx.close()
            continue
        res["flux_o1"] = time_call(ctx, stream, lambda: ctx.call("athk_calculate_fluxes", 1, dt, 0))
        res["flux_o2"] = time_call(ctx, stream, lambda: ctx.call("athk_calculate_fluxes", 2, dt, 0))
        res["corner_e"] = time_call(ctx, stream, lambda: ctx.call("athk_compute_corner_e"))
        res["emf_send"] = time_call(ctx, stream, lambda: ctx.call("athk_send_flux_correction"))
        res["emf_recv"] = time_call(ctx, stream, lambda: ctx.call("athk_recv_flux_correction"))
        res["fluxdiv"] = time_call(ctx, stream, lambda: ctx.call("athk_add_flux_divergence", 0.0, 0))
        res["ct"] = time_call(ctx, stream, lambda: ctx.call("athk_ct", 0.0, 0))
        res["ghosts_u"] = time_call(ctx, stream, lambda: ctx.call("athk_exchange_ghosts", 1, 0))
        res["ghosts_b"] = time_call(ctx, stream, lambda: ctx.call("athk_exchange_ghosts", 2, 0))
        res["c2p"] = time_call(ctx, stream, lambda: ctx.call("athk_cons2prim", 0))
        res["cycle_fused"] = time_call(ctx, stream, lambda: ti.step(), reps=3)
        ti.fused = False
        res["cycle_unfused"] = time_call(ctx, stream, lambda: ti.step(), reps=3)
        print(os.path.basename(path), " ".join("%s=%.3f" % kv for kv in res.items()), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
