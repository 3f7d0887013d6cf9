"""Golden case split across the ranks of a torchrun job (one rank per GPU): the fields every rank
ends with must match the reference's, bit for bit, whatever the transport (peer-to-peer stores or
NCCL).  Importable (bench.py runs split_golden_case() before its timed region when N > 1) and a
tool: torchrun --nproc-per-node N tools/multi_check.py [case] [--unfused]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def split_golden_case(lib, name, rank, world, local_rank, fused=True):
    """Returns a dict: case, bit_identical (this rank's blocks), cycles, links, transport."""
    import util
    from athena_b200 import multigpu, pyhost
    meta, z = util.load_golden(name)
    mesh = pyhost.UniformMesh(meta["mesh"], meta["block"], meta["xmin"], meta["xmax"], ng=meta["ng"],
                              bcs=meta.get("bcs"))
    nb = mesh.nbtotal
    per = nb // world
    parts = [list(range(r * per, (r + 1) * per if r < world - 1 else nb)) for r in range(world)]
    rank_of = {g: r for r, gl in enumerate(parts) for g in gl}
    mhd = meta["mhd"]
    ns = meta.get("nscalars") or 0
    ctx = pyhost.Context(mesh, mhd, meta["gamma"], device=local_rank, gids=parts[rank], rank_of=rank_of,
                         my_rank=rank, lib=lib, nscalars=ns,
                         rsolver=util.RSOLVERS[meta["flux"]] if meta.get("flux") else None,
                         grav_acc=meta.get("grav_acc"))
    stream = torch.cuda.ExternalStream(lib.athk_stream(ctx.h))
    xch = multigpu.attach(ctx, rank, world, device="cuda", stream=stream)
    B = (z["b1_0"], z["b2_0"], z["b3_0"]) if mhd else (None, None, None)
    blocks = util.split_to_blocks(mesh, z["u0"], *B)
    sblocks = util.split_to_blocks(mesh, z["s0"]) if ns else None
    for lb, g in enumerate(ctx.gids):
        ctx.upload("U", lb, blocks[g]["u"])
        if ns:
            ctx.upload("S", lb, sblocks[g]["u"])
        if mhd:
            ctx.upload("B_X1F", lb, blocks[g]["b1"])
            ctx.upload("B_X2F", lb, blocks[g]["b2"])
            ctx.upload("B_X3F", lb, blocks[g]["b3"])
    rmin = multigpu.reduce_min_fn("cuda")
    ti = pyhost.TimeIntegrator(ctx, meta["integrator"], meta["xorder"], meta["cfl"], fused=fused)
    ti.initialize(tlim=meta["tlim"], reduce_min=rmin)
    while ti.time < meta["tlim"]:
        ti.step(meta["tlim"], reduce_min=rmin)
    ok = ti.ncycle == meta["ncycle"] and ti.dt == meta["dt_final"]
    ng = meta["ng"]
    nx1, nx2, nx3 = mesh.block_nx
    for lb, g in enumerate(ctx.gids):
        l1, l2, l3 = mesh.loc[g]
        o1, o2, o3 = l1 * nx1, l2 * nx2, l3 * nx3
        u = ctx.download("U", lb)[:, ng:ng + nx3, ng:ng + nx2, ng:ng + nx1]
        ok = ok and np.array_equal(u, z["uf"][:, o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1])
        if ns:
            sc = ctx.download("S", lb)[:, ng:ng + nx3, ng:ng + nx2, ng:ng + nx1]
            ok = ok and np.array_equal(sc, z["sf"][:, o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1])
        if mhd:
            b2 = ctx.download("B_X2F", lb)[ng:ng + nx3, ng:ng + nx2 + 1, ng:ng + nx1]
            ok = ok and np.array_equal(b2, z["b2_f"][o3:o3 + nx3, o2:o2 + nx2 + 1, o1:o1 + nx1])
    timeouts = int(lib.athk_p2p_timeouts(ctx.h))
    res = {"case": name, "bit_identical": bool(ok and timeouts == 0), "cycles": int(ti.ncycle), "links": len(ctx.remote),
           "transport": "p2p" if xch.p2p else "nccl", "blocks": len(ctx.gids)}
    xch.close()
    ctx.close()
    # every rank must agree
    t = torch.tensor([1 if res["bit_identical"] else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    res["bit_identical"] = bool(int(t.item()))
    return res


def main():
    from athena_b200 import capi
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "mhd_linwave_vl2_plm_hlld"
    fused = "--unfused" not in sys.argv
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res = split_golden_case(capi.load(), name, rank, world, lr, fused=fused)
    print("rank %d/%d %s" % (rank, world, res), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if res["bit_identical"] else 1)


if __name__ == "__main__":
    main()
