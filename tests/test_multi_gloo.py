"""Multi-rank path on CPU: two processes (gloo), MeshBlocks split in Z-order, remote links packed
by the host build of the operator bodies (tests/emul) and exchanged with torch.distributed.  The
fields must equal the reference's golden vectors bit for bit on every rank -- i.e. the result is
independent of how blocks are distributed over ranks (SURVEY App. D / mpi_linwave.py)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, fused, q, overlap=False):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here, os.path.join(here, "emul")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import util
    from athena_b200 import multigpu, pyhost
    from emul_lib import load_emul
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = load_emul(build=False)
        meta, z = util.load_golden(name)
        mesh = pyhost.UniformMesh(meta["mesh"], meta["block"], meta["xmin"], meta["xmax"], ng=meta["ng"],
                                  bcs=meta.get("bcs"))
        ns = meta.get("nscalars") or 0
        nb = mesh.nbtotal
        per = nb // world
        parts = [list(range(r * per, (r + 1) * per if r < world - 1 else nb)) for r in range(world)]
        rank_of = {g: r for r, gl in enumerate(parts) for g in gl}
        mhd = meta["mhd"]
        ctx = pyhost.Context(mesh, mhd, meta["gamma"], gids=parts[rank], rank_of=rank_of, my_rank=rank, lib=lib,
                             nscalars=ns, rsolver=util.RSOLVERS[meta["flux"]] if meta.get("flux") else None,
                     grav_acc=meta.get("grav_acc"))
        assert len(ctx.remote) > 0
        multigpu.attach(ctx, rank, world, device="cpu")
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
        rmin = multigpu.reduce_min_fn("cpu")
        ti = pyhost.TimeIntegrator(ctx, meta["integrator"], meta["xorder"], meta["cfl"], fused=fused)
        ti.overlap = overlap   # phases 10,13,11,14,12 of athk_stage_phase instead of 0,1,2
        ti.initialize(tlim=meta["tlim"], reduce_min=rmin)
        while ti.time < meta["tlim"]:
            ti.step(meta["tlim"], reduce_min=rmin)
        assert ti.ncycle == meta["ncycle"]
        assert ti.dt == meta["dt_final"]
        ng = meta["ng"]
        nx1, nx2, nx3 = mesh.block_nx
        for lb, g in enumerate(ctx.gids):
            l1, l2, l3 = mesh.loc[g]
            o1, o2, o3 = l1 * nx1, l2 * nx2, l3 * nx3
            u = ctx.download("U", lb)[:, ng:ng + nx3, ng:ng + nx2, ng:ng + nx1]
            np.testing.assert_array_equal(u, z["uf"][:, o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1])
            if ns:
                sc = ctx.download("S", lb)[:, ng:ng + nx3, ng:ng + nx2, ng:ng + nx1]
                np.testing.assert_array_equal(sc, z["sf"][:, o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1])
            if mhd:
                b1 = ctx.download("B_X1F", lb)[ng:ng + nx3, ng:ng + nx2, ng:ng + nx1 + 1]
                np.testing.assert_array_equal(b1, z["b1_f"][o3:o3 + nx3, o2:o2 + nx2, o1:o1 + nx1 + 1])
                b3 = ctx.download("B_X3F", lb)[ng:ng + nx3 + 1, ng:ng + nx2, ng:ng + nx1]
                np.testing.assert_array_equal(b3, z["b3_f"][o3:o3 + nx3 + 1, o2:o2 + nx2, o1:o1 + nx1])
        ctx.close()
        q.put((rank, "ok"))
    except Exception as ex:   # report instead of hanging the peer
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,fused,overlap", [
    ("mhd_linwave_vl2_plm_hlld", True, False),
    ("mhd_linwave_vl2_plm_hlld", False, False),
    ("hydro_linwave_rk3_plm_hllc", True, False),
    ("mhd_linwave_alfven_vl2", True, True),                  # overlap-window phases
    ("mhd_kh_rk2_plm_hlld_scalar", True, True),              # passive scalars in the ghost messages
    ("mhd_shocktube_vl2_plm_roe_llf", True, False),          # solver choice
    ("mhd_blast_vl2_plm_hlld_reflecting_box", True, False),  # physical boundaries + remote links
    ("hydro_blast_rk3_plm_hllc_mixed_bcs", False, False),
    ("mhd_blast_rk2_plm_hlld_gravity", True, True)])         # source terms in the overlap-window phases
def test_two_ranks_match_reference_bitwise(emul_lib, name, fused, overlap):
    world = 2
    port = _free_port()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, name, fused, q, overlap)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", "rank %d: %s" % (r, msg)
