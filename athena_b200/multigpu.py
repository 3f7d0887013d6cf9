"""Inter-GPU boundary exchange for MeshBlocks partitioned across ranks (one process per GPU).

Mirrors the rank-crossing half of the reference's BoundaryVariable machinery
(/root/reference/src/bvals/bvals_var.cpp:212-262 Send/ReceiveBoundaryBuffers, cc/bvals_cc.cpp:
614-663 persistent MPI requests, fc/flux_correction_fc.cpp:623-654,1610-1749 EMF correction):
every (block, neighbour) pair whose neighbour lives on another rank is a "remote link".  Where
the reference posts one MPI message per link, all links towards the same peer are coalesced here
into ONE message per peer and quantity (ghost zones of u+b; EMF surfaces), packed by a single
kernel launch.  Transport, two forms:
  * peer-to-peer (default on GPUs): every rank's receive buffers are CUDA-IPC mapped into its
    peers, the pack kernels store the messages straight into the receiver's memory over NVLink and
    raise a flag; no NCCL call, no host involvement on the data path (include/athk.h,
    athk_remote_set_p2p).  torch.distributed only carries the 64-byte IPC handles at set-up;
  * NCCL send/recv (ATHK_P2P=0) / gloo on CPU tensors for the world_size-2 tests.
There is no collective on the data path: the only reduction is the global min of dt
(Mesh::NewTimeStep, mesh.cpp:1263).
"""
import os
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def _slot(o):
    return (o[2] + 1) * 9 + (o[1] + 1) * 3 + (o[0] + 1)


class RemoteExchange:
    def __init__(self, ctx, rank, world, device, stream=None):
        self.ctx = ctx
        self.rank = rank
        self.world = world
        self.stream = stream
        self.comm_stream = None
        self._pending = []
        lib, h = ctx.lib, ctx.h
        links = ctx.remote          # [(lb, (o1,o2,o3), remote gid, remote rank)], index = link id
        n = len(links)
        self.nlinks = n
        gcount = [lib.athk_remote_ghost_count(h, *l[1]) for l in links]
        ecount = [lib.athk_remote_emf_count(h, *l[1]) for l in links]
        peers = sorted(set(l[3] for l in links))
        self.peers = peers
        gs = np.zeros(n, dtype=np.int64)
        gr = np.zeros(n, dtype=np.int64)
        es = np.zeros(n, dtype=np.int64)
        er = np.zeros(n, dtype=np.int64)
        self.gseg, self.eseg = {}, {}   # peer -> (send start, send end, recv start, recv end)
        g_s = g_r = e_s = e_r = 0
        for p in peers:
            mine = [q for q in range(n) if links[q][3] == p]
            # canonical order of messages between two ranks: (receiver's block gid, offset as the
            # receiver sees it).  Sender side: receiver block = remote gid, offset = -ox.
            send_order = sorted(mine, key=lambda q: (links[q][2], _slot(tuple(-x for x in links[q][1]))))
            recv_order = sorted(mine, key=lambda q: (ctx.gids[links[q][0]], _slot(links[q][1])))
            gs0, gr0, es0, er0 = g_s, g_r, e_s, e_r
            for q in send_order:
                gs[q] = g_s
                g_s += gcount[q]
                es[q] = e_s
                e_s += ecount[q]
            for q in recv_order:
                gr[q] = g_r
                g_r += gcount[q]
                er[q] = e_r
                e_r += ecount[q]
            self.gseg[p] = (gs0, g_s, gr0, g_r)
            self.eseg[p] = (es0, e_s, er0, e_r)
        kw = dict(dtype=torch.float64, device=device)
        self.gsbuf = torch.zeros(max(g_s, 1), **kw)
        self.grbuf = torch.zeros(max(g_r, 1), **kw)
        self.esbuf = torch.zeros(max(e_s, 1), **kw)
        self.erbuf = torch.zeros(max(e_r, 1), **kw)
        self.ghost_bytes = 8 * g_s
        self.emf_bytes = 8 * e_s
        ia = lambda v: np.ascontiguousarray(v, dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int))   # noqa: E731
        la = lambda v: v.ctypes.data_as(C.POINTER(C.c_longlong))                                      # noqa: E731
        lb = [l[0] for l in links]
        o1 = [l[1][0] for l in links]
        o2 = [l[1][1] for l in links]
        o3 = [l[1][2] for l in links]
        self._keep = (gs, gr, es, er)
        ctx.call("athk_remote_configure", n, ia(lb), ia(o1), ia(o2), ia(o3), la(gs), la(gr), la(es), la(er))
        ctx.call("athk_remote_set_buffers", self.gsbuf.data_ptr(), self.grbuf.data_ptr(),
                 self.esbuf.data_ptr(), self.erbuf.data_ptr())
        self.p2p = False
        if stream is not None and os.environ.get("ATHK_P2P", "1") != "0":
            self._setup_p2p(links, gs, es, g_r, e_r)

    def _setup_p2p(self, links, gs, es, g_r, e_r):
        """Map every peer's receive buffers and flag words into this process (CUDA IPC) and tell the
        library where each link's message goes in the receiver's memory."""
        ctx, lib, h = self.ctx, self.ctx.lib, self.ctx.h
        n, peers = len(links), self.peers
        vp = C.c_void_p

        def alloc(nbytes):
            ptr, hd = vp(), C.create_string_buffer(64)
            ctx.call("athk_p2p_alloc", C.c_longlong(max(int(nbytes), 64)), C.byref(ptr), hd)
            return ptr.value, hd.raw

        g_ptr, g_h = alloc(2 * 8 * g_r)
        e_ptr, e_h = alloc(2 * 8 * e_r)
        f_ptr, f_h = alloc(2 * 8 * max(len(peers), 1))
        self._p2p_own = (g_ptr, e_ptr, f_ptr)
        mine = {"g": g_h, "e": e_h, "f": f_h, "g_r": int(g_r), "e_r": int(e_r), "peers": list(peers),
                "gr0": {p: int(self.gseg[p][2]) for p in peers}, "er0": {p: int(self.eseg[p][2]) for p in peers}}
        allinfo = [None] * self.world
        dist.all_gather_object(allinfo, mine)
        self._p2p_maps = []
        peer_g, peer_e, peer_f = {}, {}, {}
        for p in peers:
            for key, store in (("g", peer_g), ("e", peer_e), ("f", peer_f)):
                ptr = vp()
                ctx.call("athk_p2p_open", allinfo[p][key], C.byref(ptr))
                store[p] = ptr.value
                self._p2p_maps.append(ptr.value)
        gdst = (vp * (2 * n))()
        edst = (vp * (2 * n))()
        for q in range(n):
            p = links[q][3]
            info = allinfo[p]
            goff = info["gr0"][self.rank] + (int(gs[q]) - self.gseg[p][0])
            eoff = info["er0"][self.rank] + (int(es[q]) - self.eseg[p][0])
            for par in (0, 1):
                gdst[par * n + q] = peer_g[p] + 8 * (par * info["g_r"] + goff)
                edst[par * n + q] = peer_e[p] + 8 * (par * info["e_r"] + eoff)
        pflags = (vp * len(peers))()
        for k, p in enumerate(peers):
            pflags[k] = peer_f[p] + 16 * allinfo[p]["peers"].index(self.rank)
        ctx.call("athk_remote_set_p2p", len(peers), gdst, edst, vp(g_ptr), C.c_longlong(int(g_r)), vp(e_ptr),
                 C.c_longlong(int(e_r)), pflags, vp(f_ptr))
        dist.barrier()   # every rank has mapped its peers before the first message is written
        self.p2p = True

    def _ops(self, sbuf, rbuf, seg):
        ops = []
        for p in self.peers:
            s0, s1, r0, r1 = seg[p]
            if s1 > s0:
                ops.append(dist.P2POp(dist.isend, sbuf[s0:s1], p))
            if r1 > r0:
                ops.append(dist.P2POp(dist.irecv, rbuf[r0:r1], p))
        return ops

    # Overlapped form: start_*() orders the transfers after the pack kernels already queued on the
    # context's stream but runs them on a separate communication stream, so kernels queued on the
    # context's stream afterwards (interior corner-E, interior W(U)) execute while the messages
    # cross NVLink; finish_*() makes the context's stream wait for them before the unpack kernels.
    def _start(self, sbuf, rbuf, seg):
        ops = self._ops(sbuf, rbuf, seg) if self.peers else []
        if not ops:
            return
        if self.stream is None:          # CPU tensors (gloo tests): synchronous
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            return
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=sbuf.device)
        packed = torch.cuda.Event()
        packed.record(self.stream)
        self.comm_stream.wait_event(packed)
        with torch.cuda.stream(self.comm_stream):
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        self._pending.append(done)

    def _finish(self):
        for ev in self._pending:
            self.stream.wait_event(ev)
        self._pending = []

    def start_ghosts(self):
        if not self.p2p:
            self._start(self.gsbuf, self.grbuf, self.gseg)

    def start_emf(self):
        if self.ctx.mhd and not self.p2p:
            self._start(self.esbuf, self.erbuf, self.eseg)

    finish_ghosts = _finish
    finish_emf = _finish

    def _exchange(self, sbuf, rbuf, seg):
        ops = self._ops(sbuf, rbuf, seg) if self.peers else []
        if not ops:
            return
        if self.stream is not None:
            # NCCL work is ordered after the pack kernels on the context's stream, and the stream
            # waits for the transfers before the unpack kernels that follow
            with torch.cuda.stream(self.stream):
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        else:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def exchange_ghosts(self):
        if not self.p2p:   # (peer-to-peer: the pack kernel already delivered the messages)
            self._exchange(self.gsbuf, self.grbuf, self.gseg)

    def exchange_emf(self):
        if self.ctx.mhd and not self.p2p:
            self._exchange(self.esbuf, self.erbuf, self.eseg)

    def close(self):
        """Unmap the peers' buffers and release this rank's (after a barrier: nobody writes any more)."""
        if not self.p2p:
            return
        self.ctx.sync()
        dist.barrier()
        for ptr in self._p2p_maps:
            self.ctx.call("athk_p2p_close", C.c_void_p(ptr))
        dist.barrier()
        for ptr in self._p2p_own:
            self.ctx.call("athk_p2p_free", C.c_void_p(ptr))
        self.p2p = False


def attach(ctx, rank, world, device="cuda", stream=None):
    """Create the exchange object for a context whose neighbour table has remote links."""
    ctx.xch = RemoteExchange(ctx, rank, world, device, stream)
    return ctx.xch


def reduce_min_fn(device):
    def f(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())
    return f
