"""Several independent meshes (an ensemble, a parameter scan) advanced through HOST buffers on one GPU.

The C ABI is per context (`athk_create`): each member owns its arrays, its stream and its staging
buffer, so the copies of one member can run under the cycle of another and under the
opposite-direction copy of a third (PCIe is full duplex; copy engines and SMs are separate units).
`HostSteppedEnsemble` is the host side of that: one thread per member, every step of a member being

    pinned host (u, b) -> device,  ghost exchange + W(U),  one integrator cycle,  (u, b) -> pinned host

with one member at a time owning each direction of the link (the DMA engines interleave the copies
of concurrent streams, which drives free-running members into lockstep -- all copying in, then all
copying out -- and leaves each direction idle half of the time).  Used by `bench.py` for its `e2e`
number; there is no CPU fallback here either: every call goes to `libathk.so`."""
import ctypes as C
import threading
import time

from . import capi


class Member:
    def __init__(self, ctx, integrator, host_arrays, device=0):
        self.ctx, self.ti, self.host = ctx, integrator, host_arrays   # host: {"U": ndarray (pinned), ...}
        self.device = device

    def upload(self):
        dp = C.POINTER(C.c_double)
        for k, a in self.host.items():
            self.ctx.call("athk_upload_all", capi.ARRAYS[k], a.ctypes.data_as(dp), 1)

    def download(self):
        dp = C.POINTER(C.c_double)
        for k, a in self.host.items():
            self.ctx.call("athk_download_all_async", capi.ARRAYS[k], a.ctypes.data_as(dp), 1)
        self.ctx.sync()


class HostSteppedEnsemble:
    def __init__(self, members):
        self.members = list(members)
        self.link_in, self.link_out = threading.Lock(), threading.Lock()

    def step_member(self, m):
        with self.link_in:
            m.upload()
            m.ctx.sync()
        # the state came from the host: the start-of-cycle exchange + W(U), as after a restart
        m.ti.refresh_ghosts_and_primitives()
        m.ti.step()
        with self.link_out:
            m.download()

    def run(self, nsteps, lag_s=0.0):
        """Every member does `nsteps` steps; member n starts n * lag_s late (a fraction of a step, so
        that the members enter the pipeline one after the other instead of queueing on the first copy)."""
        errs = []

        def work(n):
            try:
                import torch
                if torch.cuda.is_available():   # (every ABI call also makes its context's device current)
                    torch.cuda.set_device(self.members[n].device)
                if lag_s > 0.0:
                    time.sleep(n * lag_s)
                for _ in range(nsteps):
                    self.step_member(self.members[n])
            except Exception as ex:   # raised in the caller after the join
                errs.append(ex)
        ths = [threading.Thread(target=work, args=(n,)) for n in range(len(self.members))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        if errs:
            raise errs[0]

    def sync(self):
        for m in self.members:
            m.ctx.sync()
