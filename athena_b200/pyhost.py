"""Python host-side mirror of the reference's Mesh / TimeIntegratorTaskList logic, driving the
C ABI (include/athk.h) directly.  It exists so that tests and bench.py can exercise the operator
boundary call by call; the C++ host (athena_b200/csrc/host) keeps the reference's class API.

Reference behaviour mirrored here (file:line under /root/reference/src):
  * block decomposition and Z-ordering     mesh/mesh.cpp:340-567, mesh/meshblock_tree.cpp:336-350
  * block extents / coordinates            mesh/mesh.cpp:1984-2060, coordinates/coordinates.cpp:100-324,
                                           mesh/mesh.hpp:457-473,535-560, coordinates/cartesian.cpp:24-60
  * neighbour search (uniform, periodic)   bvals/bvals_base.cpp:301-600
  * stage weights, task order              task_list/time_integrator.cpp:111-606, 905-1206, 1606-2318
  * global time step                       mesh/mesh.cpp:1236-1282, main.cpp:457-541
"""
import ctypes as C
import os
import sys

import numpy as np

from . import capi

_DP = C.POINTER(C.c_double)


def _dp(a):
    return a.ctypes.data_as(_DP)


def _mesh_gen_x(index, nrange):
    # ComputeMeshGeneratorX(index, nrange, sym_interval=true), mesh.hpp:457-473
    noffset = index - nrange // 2
    noffset_ceil = index - (nrange + 1) // 2
    return float(noffset + noffset_ceil) / (2.0 * nrange)


def _uniform_gen(x, xmin, xmax):
    # UniformMeshGeneratorX1, mesh.hpp:535-538
    return 0.5 * (xmin + xmax) + (x * xmax - x * xmin)


def morton_key(lx1, lx2, lx3):
    k = 0
    for bit in range(21):
        k |= ((lx1 >> bit) & 1) << (3 * bit)
        k |= ((lx2 >> bit) & 1) << (3 * bit + 1)
        k |= ((lx3 >> bit) & 1) << (3 * bit + 2)
    return k


class UniformMesh:
    """Uniform Cartesian mesh cut into equal MeshBlocks, listed in the reference's Z-order."""

    def __init__(self, mesh_nx, block_nx, xmin, xmax, ng=2, periodic=(True, True, True), bcs=None):
        """bcs: the six <mesh> flags (ix1_bc, ox1_bc, ix2_bc, ox2_bc, ix3_bc, ox3_bc) as
        'periodic' | 'reflecting' | 'outflow'; overrides ``periodic``."""
        if bcs is not None:
            bcs = tuple(bcs)
            for d in range(3):
                if (bcs[2 * d] == "periodic") != (bcs[2 * d + 1] == "periodic"):
                    raise ValueError("periodic boundaries must be paired (bvals_base.cpp)")
            periodic = tuple(bcs[2 * d] == "periodic" for d in range(3))
        else:
            bcs = tuple("periodic" for _ in range(6))
            if not all(periodic):
                raise ValueError("non-periodic mesh: give the six boundary flags (bcs=...)")
        self.bcs = bcs
        self.mesh_nx = tuple(mesh_nx)
        self.block_nx = tuple(block_nx)
        self.xmin = tuple(float(v) for v in xmin)
        self.xmax = tuple(float(v) for v in xmax)
        self.ng = ng
        self.periodic = tuple(periodic)
        for d in range(3):
            if self.mesh_nx[d] % self.block_nx[d]:
                raise ValueError("mesh nx%d not divisible by MeshBlock nx%d" % (d + 1, d + 1))
        self.nrbx = tuple(self.mesh_nx[d] // self.block_nx[d] for d in range(3))
        locs = [(a, b, c) for c in range(self.nrbx[2]) for b in range(self.nrbx[1])
                for a in range(self.nrbx[0])]
        locs.sort(key=lambda l: morton_key(*l))
        self.loc = locs
        self.nbtotal = len(locs)
        self.gid_of = {l: n for n, l in enumerate(locs)}

    def block_bcs(self, loc):
        """BoundaryValues::block_bcs as enum athk_bc: a face on the mesh edge takes the mesh flag
        unless periodic; interior faces are 'block' (0)."""
        code = {"periodic": 0, "reflecting": 1, "outflow": 2, "user": 3}
        out = []
        for d in range(3):
            out.append(code[self.bcs[2 * d]] if loc[d] == 0 else 0)
            out.append(code[self.bcs[2 * d + 1]] if loc[d] == self.nrbx[d] - 1 else 0)
        return out

    def block_bounds(self, loc, d):
        lx = loc[d]
        n = self.nrbx[d]
        lo = self.xmin[d] if lx == 0 else _uniform_gen(_mesh_gen_x(lx, n), self.xmin[d], self.xmax[d])
        hi = self.xmax[d] if lx == n - 1 else _uniform_gen(_mesh_gen_x(lx + 1, n), self.xmin[d],
                                                            self.xmax[d])
        return lo, hi

    def coords(self, loc, d):
        """(xf[nc+1], xv[nc], dxf[nc]) of a block along direction d, as Coordinates builds them."""
        nx = self.block_nx[d]
        ng = self.ng
        nc = nx + 2 * ng
        lo, hi = self.block_bounds(loc, d)
        nrootmesh = self.mesh_nx[d]
        xf = np.empty(nc + 1)
        for i in range(nc + 1):
            noffset = (i - ng) + loc[d] * nx
            xf[i] = _uniform_gen(_mesh_gen_x(noffset, nrootmesh), self.xmin[d], self.xmax[d])
        xf[ng] = lo
        xf[ng + nx] = hi
        dx = (hi - lo) / nx
        dxf = np.full(nc, dx)
        xv = 0.5 * (xf[1:] + xf[:-1])
        return xf, xv, dxf

    def neighbor_table(self, gids=None, rank_of=None, my_rank=0):
        """27-entry same-level neighbour table per local block (include/athk.h).
        gids: the global ids owned locally (default all); rank_of[gid] -> owner rank.
        Returns (table[nlocal*27] int32, remote list [(local lb, (ox1,ox2,ox3), gid, rank)])."""
        gids = list(range(self.nbtotal)) if gids is None else list(gids)
        local_of = {g: n for n, g in enumerate(gids)}
        tab = np.full((len(gids), 27), -1, dtype=np.int32)
        remote = []
        for lb, g in enumerate(gids):
            loc = self.loc[g]
            for o3 in (-1, 0, 1):
                for o2 in (-1, 0, 1):
                    for o1 in (-1, 0, 1):
                        if o1 == 0 and o2 == 0 and o3 == 0:
                            continue
                        n = [loc[0] + o1, loc[1] + o2, loc[2] + o3]
                        ok = True
                        for d in range(3):
                            if n[d] < 0 or n[d] >= self.nrbx[d]:
                                if self.periodic[d]:
                                    n[d] %= self.nrbx[d]
                                else:
                                    ok = False
                        slot = (o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1)
                        if not ok:
                            continue
                        ng_ = self.gid_of[tuple(n)]
                        if ng_ in local_of:
                            tab[lb, slot] = local_of[ng_]
                        else:
                            tab[lb, slot] = -(2 + len(remote))
                            remote.append((lb, (o1, o2, o3), ng_,
                                           rank_of[ng_] if rank_of is not None else -1))
        return tab.reshape(-1), remote


# integrator tables: (delta, gamma_1, gamma_2, gamma_3, beta) per stage; cfl_limit(ndim=3)
TWO_3RD = 0.66666666666666667
ONE_3RD = 0.33333333333333333
INTEGRATORS = {
    "vl2": ([(1.0, 0.0, 1.0, 0.0, 0.5), (0.0, 0.0, 1.0, 0.0, 1.0)], 0.5),
    "rk1": ([(1.0, 0.0, 1.0, 0.0, 1.0)], 1.0),
    "rk2": ([(1.0, 0.0, 1.0, 0.0, 1.0), (0.0, 0.5, 0.5, 0.0, 0.5)], 1.0),
    "rk3": ([(1.0, 0.0, 1.0, 0.0, 1.0), (0.0, 0.25, 0.75, 0.0, 0.25),
             (0.0, TWO_3RD, ONE_3RD, 0.0, TWO_3RD)], 1.0),
}


class Context:
    """One GPU context = the MeshBlocks of one rank."""

    def __init__(self, mesh, mhd, gamma, rsolver=None, nscalars=0, device=0, gids=None,
                 rank_of=None, my_rank=0, lib=None, grav_acc=None, arith=0, eos=0, gamma_max=0.0):
        self.lib = lib if lib is not None else capi.load()
        self.mesh = mesh
        self.mhd = bool(mhd)
        self.gids = list(range(mesh.nbtotal)) if gids is None else list(gids)
        self.nb = len(self.gids)
        self.nscalars = nscalars
        cfg = capi.Config()
        cfg.nx1, cfg.nx2, cfg.nx3 = mesh.block_nx
        cfg.nghost = mesh.ng
        cfg.nblocks = self.nb
        cfg.mhd = 1 if mhd else 0
        cfg.nscalars = nscalars
        cfg.rsolver = rsolver if rsolver is not None else (capi.RS_HLLD if mhd else capi.RS_HLLC)
        cfg.device = device
        cfg.arith = arith      # capi.ARITH_STRICT (bit-exact, default) / capi.ARITH_FAST
        cfg.gamma = gamma
        cfg.dfloor = cfg.pfloor = cfg.sfloor = 0.0
        cfg.eos = eos          # capi.EOS_ADIABATIC (default) / EOS_ISOTHERMAL needs iso_sound_speed / EOS_ADIABATIC_SR
        cfg.gamma_max = gamma_max
        h = C.c_void_p()
        capi.check(self.lib, self.lib.athk_create(C.byref(cfg), C.byref(h)))
        self.h = h
        if grav_acc is not None and any(float(x) != 0.0 for x in grav_acc):   # <hydro> grav_acc1/2/3
            capi.check(self.lib, self.lib.athk_set_constant_acceleration(h, *[float(x) for x in grav_acc]))
        self.nc = tuple(n + 2 * mesh.ng for n in mesh.block_nx)
        for lb, g in enumerate(self.gids):
            cs = []
            for d in range(3):
                cs += list(mesh.coords(mesh.loc[g], d))
            capi.check(self.lib, self.lib.athk_set_block_coords(h, lb, *[_dp(a) for a in cs]))
        tab, self.remote = mesh.neighbor_table(self.gids, rank_of, my_rank)
        self.nbr = np.ascontiguousarray(tab, dtype=np.int32)
        capi.check(self.lib, self.lib.athk_set_neighbors(h, self.nbr.ctypes.data_as(C.POINTER(C.c_int))))
        self.has_physical_bcs = False
        for lb, g in enumerate(self.gids):
            fl = mesh.block_bcs(mesh.loc[g])
            if any(fl):
                self.has_physical_bcs = True
                capi.check(self.lib, self.lib.athk_set_boundary_flags(h, lb, (C.c_int * 6)(*fl)))

    def enroll_user_boundary(self, face, fn):
        """Mesh::EnrollUserBoundaryFunction for the device path: fn(args: capi.BvalArgs) -> int is
        called once per flagged local block at the face's place in the x1, x2, x3 order and must
        enqueue its work on args.stream (athk_enroll_user_boundary)."""
        cb = capi.BVAL_FN(lambda user, a: int(fn(a.contents) or 0))
        self._user_bcs = getattr(self, "_user_bcs", []) + [cb]   # keep the thunk alive
        capi.check(self.lib, self.lib.athk_enroll_user_boundary(self.h, face, C.cast(cb, C.c_void_p), None))

    def close(self):
        if self.h:
            self.lib.athk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        capi.check(self.lib, getattr(self.lib, name)(self.h, *args))

    def shape(self, array):
        nvar, a1, a2, a3 = capi.SHAPES[array]
        if nvar is None:
            nvar = self.nscalars
        sh = (self.nc[2] + a3, self.nc[1] + a2, self.nc[0] + a1)
        multi = capi.SHAPES[array][0] != 1
        return ((nvar,) + sh) if multi else sh

    def set_block_floors(self, lb, dfloor, pfloor, sfloor):
        """Floors of local block lb (each MeshBlock of the reference owns an EquationOfState)."""
        self.call("athk_set_block_floors", lb, float(dfloor), float(pfloor), float(sfloor))

    def upload(self, array, lb, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != self.shape(array):
            raise ValueError("%s: expected shape %s, got %s" % (array, self.shape(array), a.shape))
        self.call("athk_upload", capi.ARRAYS[array], lb, _dp(a))
        self.call("athk_sync")

    def download(self, array, lb):
        a = np.empty(self.shape(array))
        self.call("athk_download", capi.ARRAYS[array], lb, _dp(a))
        return a

    def sync(self):
        self.call("athk_sync")

    def launches(self):
        return self.lib.athk_launch_count(self.h)


class TimeIntegrator:
    """Mirror of TimeIntegratorTaskList for the uniform-mesh configs (task order of
    time_integrator.cpp:905-1206 collapsed into stream order: every task acts on all blocks)."""

    def __init__(self, ctx, integrator="vl2", xorder=2, cfl_number=0.3, fused=False):
        self.ctx = ctx
        self.integrator = integrator
        self.stages, cfl_limit = INTEGRATORS[integrator]
        self.xorder = xorder
        # cfl clamp, time_integrator.cpp:893-903
        self.cfl = min(cfl_number, cfl_limit)
        self.fused = fused
        self.fold_dt = bool(int(os.environ.get("ATHK_FOLD_DT", "1")))
        self.overlap = bool(int(os.environ.get("ATHK_OVERLAP", "0")))
        self.time = 0.0
        self.dt = sys.float_info.max
        self.ncycle = 0
        self.dt_history = []

    # Mesh::Initialize after the problem generator (mesh.cpp:1686-1965)
    def initialize(self, tlim=None, reduce_min=None):
        self.refresh_ghosts_and_primitives()
        self.new_time_step(tlim, reduce_min)

    def refresh_ghosts_and_primitives(self):
        c = self.ctx
        xch = getattr(c, "xch", None)
        if xch is not None:   # neighbours on other GPUs: pack, send/recv, then fill
            c.call("athk_remote_pack_ghosts", 0)
            xch.exchange_ghosts()
        c.call("athk_exchange_ghosts", 1 | (2 if c.mhd else 0) | (4 if c.nscalars else 0), 0)
        c.call("athk_cons2prim", 0)
        c.call("athk_apply_physical_boundaries")   # mesh.cpp:1886-1891 (no-op on a periodic mesh)

    def new_block_dt(self):
        m = C.c_double()
        self.ctx.call("athk_new_block_dt", self.cfl, None, C.byref(m))
        return m.value

    def new_time_step(self, tlim, reduce_min=None):
        # Mesh::NewTimeStep, mesh.cpp:1236-1282
        blk = self.new_block_dt()
        if reduce_min is not None:
            blk = reduce_min(blk)
        dt = 2.0 * self.dt
        dt = min(dt, blk)
        if tlim is not None and self.time < tlim and (tlim - self.time) < dt:
            dt = tlim - self.time
        self.dt = dt

    def one_stage(self, stage):
        c = self.ctx
        delta, g1, g2, g3, beta = self.stages[stage - 1]
        dt = self.dt
        mhd = c.mhd
        order = 1 if (self.integrator == "vl2" and stage == 1) else self.xorder
        xch = getattr(c, "xch", None)
        if self.fused:
            sw = capi.StageWghts(delta, g1, g2, g3, beta)
            if xch is None:
                c.call("athk_stage", stage, order, C.byref(sw), dt)
            elif not self.overlap:
                c.call("athk_stage_phase", 0, stage, order, C.byref(sw), dt)
                xch.exchange_emf()
                c.call("athk_stage_phase", 1, stage, order, C.byref(sw), dt)
                xch.exchange_ghosts()
                c.call("athk_stage_phase", 2, stage, order, C.byref(sw), dt)
            else:
                # the inter-GPU transfers fly under the interior corner-E and the interior W(U)
                c.call("athk_stage_phase", 10, stage, order, C.byref(sw), dt)
                xch.start_emf()
                c.call("athk_stage_phase", 13, stage, order, C.byref(sw), dt)
                xch.finish_emf()
                c.call("athk_stage_phase", 11, stage, order, C.byref(sw), dt)
                xch.start_ghosts()
                c.call("athk_stage_phase", 14, stage, order, C.byref(sw), dt)
                xch.finish_ghosts()
                c.call("athk_stage_phase", 12, stage, order, C.byref(sw), dt)
            return
        if stage == 1:   # StartupTaskList, time_integrator.cpp:1608-1619
            c.call("athk_zero_register", 0, 1)
            if mhd:
                c.call("athk_zero_register", 1, 1)
            if c.nscalars:
                c.call("athk_zero_register", 2, 1)   # ps->s1.ZeroClear(), time_integrator.cpp:1628-1630
        # CALC_HYDFLX: donor cell on the VL2 predictor (time_integrator.cpp:1700-1701)
        order = 1 if (self.integrator == "vl2" and stage == 1) else self.xorder
        c.call("athk_calculate_fluxes", order, dt, 0)
        if c.nscalars:
            c.call("athk_calculate_scalar_fluxes", order)   # CALC_SCLRFLX (:2329-2340)
        if mhd:
            c.call("athk_compute_corner_e")          # CALC_FLDFLX
            c.call("athk_send_flux_correction")      # SEND_FLDFLX
            if xch is not None:
                c.call("athk_remote_pack_emf")
                xch.exchange_emf()
            c.call("athk_recv_flux_correction")      # RECV_FLDFLX
        # INT_HYD (time_integrator.cpp:1797-1847)
        w1 = (C.c_double * 5)(1.0, delta, 0.0, 0.0, 0.0)
        c.call("athk_weighted_ave_cc", 0, 1, 0, 2, w1)
        w2 = (C.c_double * 5)(g1, g2, g3, 0.0, 0.0)
        swap = (g1 == 0.0 and g2 == 1.0 and g3 == 0.0)
        if swap:
            c.call("athk_swap_registers", 0, 0, 1)
        else:
            c.call("athk_weighted_ave_cc", 0, 0, 1, 2, w2)
        c.call("athk_add_flux_divergence", beta * dt, 0)
        c.call("athk_add_source_terms", beta * dt)   # SRC_TERM (time_integrator.cpp:1890-1913)
        if c.nscalars:   # INT_SCLR (time_integrator.cpp:2376-2403)
            c.call("athk_weighted_ave_cc", 1, 1, 0, 2, w1)
            if swap:
                c.call("athk_swap_registers", 2, 0, 1)
            else:
                c.call("athk_weighted_ave_cc", 1, 0, 1, 2, w2)
            c.call("athk_add_scalar_flux_divergence", beta * dt, 0)
        if mhd:   # INT_FLD (time_integrator.cpp:1853-1885)
            c.call("athk_weighted_ave_fc", 1, 0, 2, w1)
            if swap:
                c.call("athk_swap_registers", 1, 0, 1)
            else:
                c.call("athk_weighted_ave_fc", 0, 1, 2, w2)
            c.call("athk_ct", beta * dt, 0)
        # SEND/RECV/SETB_HYD, _FLD ; CONS2PRIM
        if xch is not None:
            c.call("athk_remote_pack_ghosts", 0)
            xch.exchange_ghosts()
        c.call("athk_exchange_ghosts", 1 | (2 if mhd else 0) | (4 if c.nscalars else 0), 0)
        c.call("athk_cons2prim", 0)
        c.call("athk_apply_physical_boundaries")   # PHY_BVAL (time_integrator.cpp:2276-2311)

    def step(self, tlim=None, reduce_min=None):
        nst = len(self.stages)
        for stage in range(1, nst + 1):
            if stage == nst and self.fold_dt:
                # NEW_DT folded into the last stage's W(U) (same per-cell values, exact min)
                self.ctx.call("athk_request_block_dt")
            self.one_stage(stage)
        self.ncycle += 1
        self.time += self.dt
        self.dt_history.append(self.dt)
        self.new_time_step(tlim, reduce_min)

    def run(self, tlim, nlim=-1):
        while self.time < tlim and (nlim < 0 or self.ncycle < nlim):
            self.step(tlim)
