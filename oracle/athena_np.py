"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the reference's per-MeshBlock
finite-volume update (Athena++ v24.0, 3-D Cartesian uniform periodic mesh, adiabatic EOS).

Independent of the CUDA sources: written from the reference alone, operation for operation
(same association order, no FMA: numpy evaluates each binary op separately in IEEE double), so
its results are expected to be BIT-IDENTICAL to the reference binary.  It is pinned against the
golden vectors in tests/golden (outputs of the unmodified reference, tests/test_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

It works on the GLOBAL mesh with periodic wrap (np.roll): the reference result on a uniform
periodic mesh does not depend on the MeshBlock decomposition.

Reference sources followed (file:line under /root/reference/src):
  eos/adiabatic_hydro.cpp:38-80, eos/adiabatic_mhd.cpp:41-90,149-156   cons2prim, sound/fast speed
  field/field.cpp:112-180                                              cell-centred field
  reconstruct/dc.cpp:24-110, reconstruct/plm.cpp:27-175                donor cell, PLM (van Leer)
  reconstruct/ppm.cpp:44-342, reconstruct/reconstruction.cpp:486-505   PPM (uniform Cartesian, CS limiter)
  scalars/calculate_scalar_fluxes.cpp:41-417, add_scalar_flux_divergence.cpp:42-96,
  eos/eos_scalars.cpp:31-60,161-173                                    passive scalars
  hydro/rsolvers/hydro/hllc.cpp:32-179, hydro/rsolvers/mhd/hlld.cpp:36-376
  hydro/rsolvers/hydro/{hlle,llf,roe}.cpp, hydro/rsolvers/mhd/{hlle_mhd,llf_mhd}.cpp   alternative solvers
  hydro/hydro.cpp:154-158                                              CT weight
  hydro/calculate_fluxes.cpp:38-415                                    sweeps, rotations
  field/calculate_corner_e.cpp:130-229, field/ct.cpp:31-116            corner EMF, CT
  bvals/fc/flux_correction_fc.cpp:689-760,1355-1541                    EMF correction (averaging)
  hydro/add_flux_divergence.cpp:39-96, mesh/weighted_ave.cpp:35-230    update
  hydro/new_blockdt.cpp:45-230, mesh/mesh.cpp:1236-1282                time step
  bvals/bvals.cpp:453-660, bvals/cc/*reflect*/outflow*, bvals/fc/*     physical boundaries
  hydro/srcterms/constant_acc.cpp:25-76, task_list/time_integrator.cpp:1890-1913   constant accelerations
  task_list/time_integrator.cpp:111-606,1694-1885                      stage weights / order

Floors: one value per run (eos.hpp:161).  The reference's MeshBlocks after the first actually read
their floors back from a 6-digit string ("Bit-exactness rules"); no golden fixture
lets a floor fire, that case is pinned through the host-binary tests instead.
"""
import sys

import numpy as np

IDN, IM1, IM2, IM3, IEN = 0, 1, 2, 3, 4
IVX, IVY, IVZ, IPR, IBY, IBZ = 1, 2, 3, 4, 5, 6
TINY_NUMBER = 1.0e-20
FLOOR = np.sqrt(1024 * np.float64(np.finfo(np.float32).tiny))   # eos.hpp:161
TWO_3RD = 0.66666666666666667
ONE_3RD = 0.33333333333333333
INTEGRATORS = {   # (delta, gamma_1, gamma_2, gamma_3, beta)
    "vl2": [(1.0, 0.0, 1.0, 0.0, 0.5), (0.0, 0.0, 1.0, 0.0, 1.0)],
    "rk2": [(1.0, 0.0, 1.0, 0.0, 1.0), (0.0, 0.5, 0.5, 0.0, 0.5)],
    "rk3": [(1.0, 0.0, 1.0, 0.0, 1.0), (0.0, 0.25, 0.75, 0.0, 0.25),
            (0.0, TWO_3RD, ONE_3RD, 0.0, TWO_3RD)],
}
CFL_LIMIT = {"vl2": 0.5, "rk2": 1.0, "rk3": 1.0}

AX = {0: -1, 1: -2, 2: -3}   # sweep direction -> numpy axis of a (k,j,i) array


def sh(a, d, n):
    """value at index+n along direction d (periodic)."""
    return np.roll(a, -n, axis=AX[d])


def mesh_gen_x(index, nrange):   # mesh.hpp:457-473 (sym_interval = true)
    noffset = index - nrange // 2
    noffset_ceil = index - (nrange + 1) // 2
    return float(noffset + noffset_ceil) / (2.0 * nrange)


def uniform_gen(x, xmin, xmax):   # mesh.hpp:535-538
    return 0.5 * (xmin + xmax) + (x * xmax - x * xmin)


NG = 2


class Mesh:
    """Global uniform mesh treated as ONE MeshBlock with NG ghost zones; 1-D coordinate arrays as
    Coordinates builds them (coordinates.cpp:100-324, cartesian.cpp:24-60).  Every 3-D array of
    the oracle lives in the same padded box (nc3+1, nc2+1, nc1+1) so that cell, face and edge
    data share one index space (index ng = first active cell / face)."""

    def __init__(self, nx, xmin, xmax, ng=NG, bcs=None):
        self.nx = tuple(nx)
        self.ng = ng
        # physical boundaries: per direction (inner, outer) in {None, "reflecting", "outflow"}
        bcs = list(bcs) if bcs is not None else ["periodic"] * 6
        self.bc = [tuple(None if f == "periodic" else f for f in bcs[2 * d:2 * d + 2]) for d in range(3)]
        self.nc = tuple(n + 2 * ng for n in nx)
        self.shape = tuple(n + 1 for n in self.nc[::-1])   # (k,j,i) padded box
        self.dxf, self.cp, self.cm = [], [], []
        for d in range(3):
            n, nc = nx[d], self.nc[d]
            xf = np.array([uniform_gen(mesh_gen_x(i - ng, n), xmin[d], xmax[d]) for i in range(nc + 1)])
            xf[ng] = xmin[d]
            xf[ng + n] = xmax[d]
            dx = (xmax[d] - xmin[d]) / n
            dxf = np.full(nc + 1, dx)
            xv = 0.5 * (xf[1:] + xf[:-1])
            cp = np.ones(nc + 1)
            cm = np.ones(nc + 1)
            cp[:nc] = (xf[1:] - xv) / dxf[:nc]    # plm.cpp:152
            cm[:nc] = (xv - xf[:-1]) / dxf[:nc]   # plm.cpp:153
            self.dxf.append(dxf)
            self.cp.append(cp)
            self.cm.append(cm)

    def along(self, arr1d, d):
        shape = [1, 1, 1]
        shape[AX[d]] = -1
        return arr1d.reshape(shape)

    def idx(self, d):
        """index array along direction d, broadcastable over the padded box"""
        return self.along(np.arange(self.nc[d] + 1), d)

    def active(self, d, extra=0):
        """mask of active indices [ng, ng+n-1+extra] along d"""
        i = self.idx(d)
        return (i >= self.ng) & (i <= self.ng + self.nx[d] - 1 + extra)

    def pad_cc(self, a):
        """global interior (…,n3,n2,n1) -> padded box with zero ghosts"""
        ng = self.ng
        out = np.zeros(a.shape[:-3] + self.shape)
        out[..., ng:ng + self.nx[2], ng:ng + self.nx[1], ng:ng + self.nx[0]] = a
        return out

    def pad_fc(self, a, d):
        ng = self.ng
        n = [self.nx[2], self.nx[1], self.nx[0]]
        n[2 - d] += 1
        out = np.zeros(self.shape)
        out[ng:ng + n[0], ng:ng + n[1], ng:ng + n[2]] = a
        return out

    def unpad(self, a, d=None):
        ng = self.ng
        n = [self.nx[2], self.nx[1], self.nx[0]]
        if d is not None:
            n[2 - d] += 1
        return a[..., ng:ng + n[0], ng:ng + n[1], ng:ng + n[2]]

    def fill_ghosts_cc(self, a):
        """periodic self-exchange of a cell-centred array (bvals_cc.cpp:287-302,386-420)"""
        for d in range(3):
            if self.bc[d][0] is not None:   # physical boundary: ghosts are set by apply_physical_bcs
                continue
            i = self.idx(d)
            n, ng = self.nx[d], self.ng
            lo = np.roll(a, n, axis=AX[d])    # value at index - n ... placed at index
            hi = np.roll(a, -n, axis=AX[d])
            a = np.where(i < ng, hi, np.where(i > ng + n - 1, lo, a))
        return a

    def fill_ghosts_fc(self, a, comp):
        """face field: along its own direction the shared faces ng and ng+n are kept
        (bvals_fc.cpp:591-593)"""
        for d in range(3):
            if self.bc[d][0] is not None:
                continue
            i = self.idx(d)
            n, ng = self.nx[d], self.ng
            top = ng + n - 1 + (1 if d == comp else 0)
            lo = np.roll(a, n, axis=AX[d])
            hi = np.roll(a, -n, axis=AX[d])
            a = np.where(i < ng, hi, np.where(i > top, lo, a))
        return a


def cell_centered_b(b1, b2, b3):   # field.cpp:139-170, uniform
    return (0.5 * b1 + 0.5 * sh(b1, 0, 1), 0.5 * b2 + 0.5 * sh(b2, 1, 1), 0.5 * b3 + 0.5 * sh(b3, 2, 1))


def cons2prim(u, gamma, bcc=None, dfloor=FLOOR, pfloor=FLOOR):
    """Returns w; u is modified in place where floors fire (adiabatic_*.cpp)."""
    gm1 = gamma - 1.0
    u[IDN] = np.where(u[IDN] > dfloor, u[IDN], dfloor)
    w = np.empty_like(u)
    w[IDN] = u[IDN]
    di = 1.0 / u[IDN]
    w[IVX] = u[IM1] * di
    w[IVY] = u[IM2] * di
    w[IVZ] = u[IM3] * di
    e_k = 0.5 * di * (u[IM1] * u[IM1] + u[IM2] * u[IM2] + u[IM3] * u[IM3])
    if bcc is not None:
        pb = 0.5 * (bcc[0] * bcc[0] + bcc[1] * bcc[1] + bcc[2] * bcc[2])
        p = gm1 * (u[IEN] - e_k - pb)
        u[IEN] = np.where(p > pfloor, u[IEN], (pfloor / gm1) + e_k + pb)
    else:
        p = gm1 * (u[IEN] - e_k)
        u[IEN] = np.where(p > pfloor, u[IEN], (pfloor / gm1) + e_k)
    w[IPR] = np.where(p > pfloor, p, pfloor)
    return w


def fast_speed(gamma, d, p, by, bz, bx):   # adiabatic_mhd.cpp:149-156
    asq = gamma * p
    vaxsq = bx * bx
    ct2 = (by * by + bz * bz)
    qsq = vaxsq + ct2 + asq
    tmp = vaxsq + ct2 - asq
    return np.sqrt(0.5 * (qsq + np.sqrt(tmp * tmp + 4.0 * asq * ct2)) / d)


def sign(x):   # SIGN() of defs.hpp.in:150
    return np.where(x < 0.0, -1.0, 1.0)


def ppm_cell(q, d):
    """PPM of one variable in every cell: (ql_iph, qr_imh) = left state of the cell's upper face,
    right state of its lower face.  Uniform Cartesian branch of ppm.cpp:107-320 with the weights
    of reconstruction.cpp:494-505 and the extremum-preserving CS limiter (C2 = 1.25)."""
    C2 = 1.25
    c1 = c2 = c3 = c4 = 0.5
    c5, c6 = 1.0 / 6.0, -1.0 / 6.0
    q_im2, q_im1, q_ip1, q_ip2 = sh(q, d, -2), sh(q, d, -1), sh(q, d, 1), sh(q, d, 2)
    qa = (q - q_im1)
    qb = (q_ip1 - q)
    dd_im1 = c1 * qa + c2 * (q_im1 - q_im2)
    dd = c1 * qb + c2 * qa
    dd_ip1 = c1 * (q_ip2 - q_ip1) + c2 * qb
    dph = (c3 * q_im1 + c4 * q) + (c5 * dd_im1 + c6 * dd)                       # ppm.cpp:124-127
    dph_ip1 = (c3 * q + c4 * q_ip1) + (c5 * dd + c6 * dd_ip1)
    d2qc_im1 = q_im2 + q - 2.0 * q_im1                                         # :142-146
    d2qc = q_im1 + q_ip1 - 2.0 * q
    d2qc_ip1 = q + q_ip2 - 2.0 * q_ip1

    def face_limit(dph_x, q_lo, q_hi, qb_, qc_):                               # :149-189
        qa_tmp = dph_x - q_lo
        qb_tmp = q_hi - dph_x
        qa_ = 3.0 * (q_lo + q_hi - 2.0 * dph_x)
        same = (sign(qa_) == sign(qb_)) & (sign(qa_) == sign(qc_))
        qd = np.where(same, sign(qa_) * np.minimum(C2 * np.abs(qb_), np.minimum(C2 * np.abs(qc_), np.abs(qa_))), 0.0)
        tmp = 0.5 * (q_lo + q_hi) - qd / 6.0
        return np.where(qa_tmp * qb_tmp < 0.0, tmp, dph_x)

    dph = face_limit(dph, q_im1, q, d2qc_im1, d2qc)
    dph_ip1 = face_limit(dph_ip1, q, q_ip1, d2qc, d2qc_ip1)
    d2qf = 6.0 * (dph + dph_ip1 - 2.0 * q)                                     # :192-194
    qminus, qplus = dph, dph_ip1
    dqf_minus = q - qminus                                                      # :218-222
    dqf_plus = qplus - q
    qa_tmp = dqf_minus * dqf_plus                                               # :226-279
    qb_tmp = (q_ip1 - q) * (q - q_im1)
    qa, qb, qc, qd = d2qc_im1, d2qc, d2qc_ip1, d2qf
    same = (sign(qa) == sign(qb)) & (sign(qa) == sign(qc)) & (sign(qa) == sign(qd))
    qe = np.where(same, sign(qd) * np.minimum(np.minimum(C2 * np.abs(qa), C2 * np.abs(qb)),
                                               np.minimum(C2 * np.abs(qc), np.abs(qd))), 0.0)
    qa = np.maximum(np.abs(q_im1), np.abs(q_im2))
    qb = np.maximum(np.maximum(np.abs(q), np.abs(q_ip1)), np.abs(q_ip2))
    rho = np.where(np.abs(qd) > (1.0e-12) * np.maximum(qa, qb), qe / qd, 0.0)
    tmp_m = q - rho * dqf_minus
    tmp_p = q + rho * dqf_plus
    tmp2_m = q - 2.0 * dqf_plus
    tmp2_p = q + 2.0 * dqf_minus
    extremum = (qa_tmp <= 0.0) | (qb_tmp <= 0.0)
    smooth = rho <= (1.0 - (1.0e-12))
    over_m = np.abs(dqf_minus) >= 2.0 * np.abs(dqf_plus)
    over_p = np.abs(dqf_plus) >= 2.0 * np.abs(dqf_minus)
    qminus_n = np.where(extremum, np.where(smooth, tmp_m, qminus), np.where(over_m, tmp2_m, qminus))
    qplus_n = np.where(extremum, np.where(smooth, tmp_p, qplus), np.where(over_p, tmp2_p, qplus))
    return qplus_n, qminus_n


def reconstruct(q, d, order, mesh):
    """q: list of (k,j,i) arrays. Returns (wl, wr) at face n (between cells n-1 and n)."""
    wl, wr = [], []
    cp = mesh.along(mesh.cp[d], d)
    cm = mesh.along(mesh.cm[d], d)
    with np.errstate(all="ignore"):
        for a in q:
            if order == 1:
                up, lo = a, a
            elif order == 3:
                up, lo = ppm_cell(a, d)
            else:
                dwl = a - sh(a, d, -1)
                dwr = sh(a, d, 1) - a
                dw2 = dwl * dwr
                dwm = 2.0 * dw2 / (dwl + dwr)
                dwm = np.where(dw2 <= 0.0, 0.0, dwm)
                up = a + cp * dwm      # L state of face n+1
                lo = a - cm * dwm      # R state of face n
            wl.append(sh(up, d, -1))
            wr.append(lo)
    return wl, wr


def hllc(wl, wr, gamma):   # hllc.cpp:53-175
    gm1 = gamma - 1.0
    igm1 = 1.0 / gm1
    with np.errstate(all="ignore"):
        cl = np.sqrt(gamma * wl[IPR] / wl[IDN])
        cr = np.sqrt(gamma * wr[IPR] / wr[IDN])
        el = wl[IPR] * igm1 + 0.5 * wl[IDN] * (wl[IVX] * wl[IVX] + wl[IVY] * wl[IVY] + wl[IVZ] * wl[IVZ])
        er = wr[IPR] * igm1 + 0.5 * wr[IDN] * (wr[IVX] * wr[IVX] + wr[IVY] * wr[IVY] + wr[IVZ] * wr[IVZ])
        rhoa = .5 * (wl[IDN] + wr[IDN])
        ca = .5 * (cl + cr)
        pmid = .5 * (wl[IPR] + wr[IPR] + (wl[IVX] - wr[IVX]) * rhoa * ca)
        ql = np.where(pmid <= wl[IPR], 1.0, np.sqrt(1.0 + (gamma + 1) / (2 * gamma) * (pmid / wl[IPR] - 1.0)))
        qr = np.where(pmid <= wr[IPR], 1.0, np.sqrt(1.0 + (gamma + 1) / (2 * gamma) * (pmid / wr[IPR] - 1.0)))
        al = wl[IVX] - cl * ql
        ar = wr[IVX] + cr * qr
        bp = np.where(ar > 0.0, ar, TINY_NUMBER)
        bm = np.where(al < 0.0, al, -TINY_NUMBER)
        vxl = wl[IVX] - al
        vxr = wr[IVX] - ar
        tl = wl[IPR] + vxl * wl[IDN] * wl[IVX]
        tr = wr[IPR] + vxr * wr[IDN] * wr[IVX]
        ml = wl[IDN] * vxl
        mr = -(wr[IDN] * vxr)
        am = (tl - tr) / (ml + mr)
        cp = (ml * tr + mr * tl) / (ml + mr)
        cp = np.where(cp > 0.0, cp, 0.0)
        vxl = wl[IVX] - bm
        vxr = wr[IVX] - bp
        fl = [wl[IDN] * vxl, wl[IDN] * wl[IVX] * vxl + wl[IPR], wl[IDN] * wl[IVY] * vxl,
              wl[IDN] * wl[IVZ] * vxl, el * vxl + wl[IPR] * wl[IVX]]
        fr = [wr[IDN] * vxr, wr[IDN] * wr[IVX] * vxr + wr[IPR], wr[IDN] * wr[IVY] * vxr,
              wr[IDN] * wr[IVZ] * vxr, er * vxr + wr[IPR] * wr[IVX]]
        pos = am >= 0.0
        sl = np.where(pos, am / (am - bm), 0.0)
        sr = np.where(pos, 0.0, -am / (bp - am))
        sm = np.where(pos, -bm / (am - bm), bp / (bp - am))
        f = [sl * fl[n] + sr * fr[n] for n in range(5)]
        f[IVX] = sl * fl[IVX] + sr * fr[IVX] + sm * cp
        f[IEN] = sl * fl[IEN] + sr * fr[IEN] + sm * cp * am
    return f


def _sq(x):
    return x * x


def hlle_hydro(wl, wr, gamma):   # hydro/rsolvers/hydro/hlle.cpp:53-160
    gm1 = gamma - 1.0
    igm1 = 1.0 / gm1
    with np.errstate(all="ignore"):
        sqrtdl, sqrtdr = np.sqrt(wl[IDN]), np.sqrt(wr[IDN])
        isdlpdr = 1.0 / (sqrtdl + sqrtdr)
        vroe = [(sqrtdl * wl[v] + sqrtdr * wr[v]) * isdlpdr for v in (IVX, IVY, IVZ)]
        el = wl[IPR] * igm1 + 0.5 * wl[IDN] * (_sq(wl[IVX]) + _sq(wl[IVY]) + _sq(wl[IVZ]))
        er = wr[IPR] * igm1 + 0.5 * wr[IDN] * (_sq(wr[IVX]) + _sq(wr[IVY]) + _sq(wr[IVZ]))
        hroe = ((el + wl[IPR]) / sqrtdl + (er + wr[IPR]) / sqrtdr) * isdlpdr
        cl = np.sqrt(gamma * wl[IPR] / wl[IDN])
        cr = np.sqrt(gamma * wr[IPR] / wr[IDN])
        q = hroe - 0.5 * (_sq(vroe[0]) + _sq(vroe[1]) + _sq(vroe[2]))
        a = np.where(q < 0.0, 0.0, np.sqrt(gm1 * q))
        al = np.minimum((vroe[0] - a), (wl[IVX] - cl))
        ar = np.maximum((vroe[0] + a), (wr[IVX] + cr))
        bp = np.where(ar > 0.0, ar, 0.0)
        bm = np.where(al < 0.0, al, 0.0)
        vxl = wl[IVX] - bm
        vxr = wr[IVX] - bp
        fl = [wl[IDN] * vxl, wl[IDN] * wl[IVX] * vxl + wl[IPR], wl[IDN] * wl[IVY] * vxl,
              wl[IDN] * wl[IVZ] * vxl, el * vxl + wl[IPR] * wl[IVX]]
        fr = [wr[IDN] * vxr, wr[IDN] * wr[IVX] * vxr + wr[IPR], wr[IDN] * wr[IVY] * vxr,
              wr[IDN] * wr[IVZ] * vxr, er * vxr + wr[IPR] * wr[IVX]]
        tmp = np.where(bp != bm, 0.5 * (bp + bm) / (bp - bm), 0.0)
        return [0.5 * (fl[n] + fr[n]) + (fl[n] - fr[n]) * tmp for n in range(5)]


def _hydro_lr(wl, wr, gamma):
    """physical fluxes and conserved jump shared by llf.cpp:72-112 and roe.cpp:86-121"""
    gm1 = gamma - 1.0
    mxl, mxr = wl[IDN] * wl[IVX], wr[IDN] * wr[IVX]
    el = wl[IPR] / gm1 + 0.5 * wl[IDN] * (_sq(wl[IVX]) + _sq(wl[IVY]) + _sq(wl[IVZ]))
    er = wr[IPR] / gm1 + 0.5 * wr[IDN] * (_sq(wr[IVX]) + _sq(wr[IVY]) + _sq(wr[IVZ]))
    fl = [mxl, mxl * wl[IVX] + wl[IPR], mxl * wl[IVY], mxl * wl[IVZ], (el + wl[IPR]) * wl[IVX]]
    fr = [mxr, mxr * wr[IVX] + wr[IPR], mxr * wr[IVY], mxr * wr[IVZ], (er + wr[IPR]) * wr[IVX]]
    du = [wr[IDN] - wl[IDN], wr[IDN] * wr[IVX] - wl[IDN] * wl[IVX], wr[IDN] * wr[IVY] - wl[IDN] * wl[IVY],
          wr[IDN] * wr[IVZ] - wl[IDN] * wl[IVZ], er - el]
    return fl, fr, du, el, er


def llf_hydro(wl, wr, gamma):   # hydro/rsolvers/hydro/llf.cpp:58-121
    with np.errstate(all="ignore"):
        cl = np.sqrt(gamma * wl[IPR] / wl[IDN])
        cr = np.sqrt(gamma * wr[IPR] / wr[IDN])
        a = 0.5 * np.maximum((np.abs(wl[IVX]) + cl), (np.abs(wr[IVX]) + cr))
        fl, fr, du, _, _ = _hydro_lr(wl, wr, gamma)
        return [0.5 * (fl[n] + fr[n]) - a * du[n] for n in range(5)]


def roe_hydro(wl, wr, gamma):   # hydro/rsolvers/hydro/roe.cpp:72-175 + RoeFlux :213-300
    gm1 = gamma - 1.0
    with np.errstate(all="ignore"):
        sqrtdl, sqrtdr = np.sqrt(wl[IDN]), np.sqrt(wr[IDN])
        isdlpdr = 1.0 / (sqrtdl + sqrtdr)
        v1, v2, v3 = [(sqrtdl * wl[v] + sqrtdr * wr[v]) * isdlpdr for v in (IVX, IVY, IVZ)]
        fl, fr, du, el, er = _hydro_lr(wl, wr, gamma)
        h = ((el + wl[IPR]) / sqrtdl + (er + wr[IPR]) / sqrtdr) * isdlpdr
        flx = [0.5 * (fl[n] + fr[n]) for n in range(5)]
        vsq = v1 * v1 + v2 * v2 + v3 * v3
        q = h - 0.5 * vsq
        cs_sq = np.where(q < 0.0, TINY_NUMBER, gm1 * q)
        cs = np.sqrt(cs_sq)
        ev = [v1 - cs, v1, v1, v1, v1 + cs]
        na = 0.5 / cs_sq
        a0 = du[0] * (0.5 * gm1 * vsq + v1 * cs)
        a0 = a0 - du[1] * (gm1 * v1 + cs)
        a0 = a0 - du[2] * gm1 * v2
        a0 = a0 - du[3] * gm1 * v3
        a0 = a0 + du[4] * gm1
        a0 = a0 * na
        a1 = du[0] * (-v2)
        a1 = a1 + du[2]
        a2 = du[0] * (-v3)
        a2 = a2 + du[3]
        qa = gm1 / cs_sq
        a3 = du[0] * (1.0 - na * gm1 * vsq)
        a3 = a3 + du[1] * qa * v1
        a3 = a3 + du[2] * qa * v2
        a3 = a3 + du[3] * qa * v3
        a3 = a3 - du[4] * qa
        a4 = du[0] * (0.5 * gm1 * vsq - v1 * cs)
        a4 = a4 - du[1] * (gm1 * v1 - cs)
        a4 = a4 - du[2] * gm1 * v2
        a4 = a4 - du[3] * gm1 * v3
        a4 = a4 + du[4] * gm1
        a4 = a4 * na
        a = [a0, a1, a2, a3, a4]
        coeff = [-0.5 * np.abs(ev[n]) * a[n] for n in range(5)]
        dens = wl[IDN] + a[0]
        llf = dens < 0.0
        dens = dens + a[3]
        llf = llf | (dens < 0.0)
        flx[0] = ((flx[0] + coeff[0]) + coeff[3]) + coeff[4]
        flx[1] = ((flx[1] + coeff[0] * (v1 - cs)) + coeff[3] * v1) + coeff[4] * (v1 + cs)
        flx[2] = (((flx[2] + coeff[0] * v2) + coeff[1]) + coeff[3] * v2) + coeff[4] * v2
        flx[3] = (((flx[3] + coeff[0] * v3) + coeff[2]) + coeff[3] * v3) + coeff[4] * v3
        flx[4] = ((((flx[4] + coeff[0] * (h - v1 * cs)) + coeff[1] * v2) + coeff[2] * v3)
                  + coeff[3] * 0.5 * vsq) + coeff[4] * (h + v1 * cs)
        flx = [np.where(ev[0] >= 0.0, fl[n], flx[n]) for n in range(5)]
        flx = [np.where(ev[4] <= 0.0, fr[n], flx[n]) for n in range(5)]
        cl = np.sqrt(gamma * wl[IPR] / wl[IDN])
        cr = np.sqrt(gamma * wr[IPR] / wr[IDN])
        al = 0.5 * np.maximum((np.abs(wl[IVX]) + cl), (np.abs(wr[IVX]) + cr))
        return [np.where(llf, 0.5 * (fl[n] + fr[n]) - al * du[n], flx[n]) for n in range(5)]


def _mhd_lr(wl, wr, bxi, gamma):
    """llf_mhd.cpp:78-125 == roe_mhd.cpp:108-160: physical fluxes, conserved jump"""
    gm1 = gamma - 1.0
    mxl, mxr = wl[IDN] * wl[IVX], wr[IDN] * wr[IVX]
    pbl = 0.5 * (bxi * bxi + _sq(wl[IBY]) + _sq(wl[IBZ]))
    pbr = 0.5 * (bxi * bxi + _sq(wr[IBY]) + _sq(wr[IBZ]))
    el = wl[IPR] / gm1 + 0.5 * wl[IDN] * (_sq(wl[IVX]) + _sq(wl[IVY]) + _sq(wl[IVZ])) + pbl
    er = wr[IPR] / gm1 + 0.5 * wr[IDN] * (_sq(wr[IVX]) + _sq(wr[IVY]) + _sq(wr[IVZ])) + pbr
    fl = [mxl, mxl * wl[IVX] + pbl - _sq(bxi) + wl[IPR], mxl * wl[IVY] - bxi * wl[IBY], mxl * wl[IVZ] - bxi * wl[IBZ],
          (el + wl[IPR] + pbl - bxi * bxi) * wl[IVX] - bxi * (wl[IBY] * wl[IVY] + wl[IBZ] * wl[IVZ]),
          wl[IBY] * wl[IVX] - bxi * wl[IVY], wl[IBZ] * wl[IVX] - bxi * wl[IVZ]]
    fr = [mxr, mxr * wr[IVX] + pbr - _sq(bxi) + wr[IPR], mxr * wr[IVY] - bxi * wr[IBY], mxr * wr[IVZ] - bxi * wr[IBZ],
          (er + wr[IPR] + pbr - bxi * bxi) * wr[IVX] - bxi * (wr[IBY] * wr[IVY] + wr[IBZ] * wr[IVZ]),
          wr[IBY] * wr[IVX] - bxi * wr[IVY], wr[IBZ] * wr[IVX] - bxi * wr[IVZ]]
    du = [wr[IDN] - wl[IDN], wr[IDN] * wr[IVX] - wl[IDN] * wl[IVX], wr[IDN] * wr[IVY] - wl[IDN] * wl[IVY],
          wr[IDN] * wr[IVZ] - wl[IDN] * wl[IVZ], er - el, wr[IBY] - wl[IBY], wr[IBZ] - wl[IBZ]]
    return fl, fr, du


def llf_mhd(wl, wr, bxi, gamma):   # hydro/rsolvers/mhd/llf_mhd.cpp:72-148
    with np.errstate(all="ignore"):
        cfl = fast_speed(gamma, wl[IDN], wl[IPR], wl[IBY], wl[IBZ], bxi)
        cfr = fast_speed(gamma, wr[IDN], wr[IPR], wr[IBY], wr[IBZ], bxi)
        a = 0.5 * np.maximum((np.abs(wl[IVX]) + cfl), (np.abs(wr[IVX]) + cfr))
        fl, fr, du = _mhd_lr(wl, wr, bxi, gamma)
        return [0.5 * (fl[n] + fr[n]) - a * du[n] for n in range(7)]


def hlle_mhd(wl, wr, bxi, gamma):   # hydro/rsolvers/mhd/hlle_mhd.cpp:63-170
    gm1 = gamma - 1.0
    with np.errstate(all="ignore"):
        sqrtdl, sqrtdr = np.sqrt(wl[IDN]), np.sqrt(wr[IDN])
        isdlpdr = 1.0 / (sqrtdl + sqrtdr)
        d = sqrtdl * sqrtdr
        v1, v2, v3 = [(sqrtdl * wl[v] + sqrtdr * wr[v]) * isdlpdr for v in (IVX, IVY, IVZ)]
        b2 = (sqrtdr * wl[IBY] + sqrtdl * wr[IBY]) * isdlpdr
        b3 = (sqrtdr * wl[IBZ] + sqrtdl * wr[IBZ]) * isdlpdr
        x = 0.5 * (_sq(wl[IBY] - wr[IBY]) + _sq(wl[IBZ] - wr[IBZ])) / (_sq(sqrtdl + sqrtdr))
        y = 0.5 * (wl[IDN] + wr[IDN]) / d
        pbl = 0.5 * (bxi * bxi + _sq(wl[IBY]) + _sq(wl[IBZ]))
        pbr = 0.5 * (bxi * bxi + _sq(wr[IBY]) + _sq(wr[IBZ]))
        el = wl[IPR] / gm1 + 0.5 * wl[IDN] * (_sq(wl[IVX]) + _sq(wl[IVY]) + _sq(wl[IVZ])) + pbl
        er = wr[IPR] / gm1 + 0.5 * wr[IDN] * (_sq(wr[IVX]) + _sq(wr[IVY]) + _sq(wr[IVZ])) + pbr
        hroe = ((el + wl[IPR] + pbl) / sqrtdl + (er + wr[IPR] + pbr) / sqrtdr) * isdlpdr
        cl = fast_speed(gamma, wl[IDN], wl[IPR], wl[IBY], wl[IBZ], bxi)
        cr = fast_speed(gamma, wr[IDN], wr[IPR], wr[IBY], wr[IBZ], bxi)
        btsq = _sq(b2) + _sq(b3)
        vaxsq = bxi * bxi / d
        bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq
        hp = hroe - (vaxsq + btsq / d)
        vsq = _sq(v1) + _sq(v2) + _sq(v3)
        twid_asq = np.maximum((gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x), 0.0)
        ct2 = bt_starsq / d
        tsum = vaxsq + ct2 + twid_asq
        tdif = vaxsq + ct2 - twid_asq
        cf2_cs2 = np.sqrt(tdif * tdif + 4.0 * twid_asq * ct2)
        a = np.sqrt(0.5 * (tsum + cf2_cs2))
        al = np.minimum((v1 - a), (wl[IVX] - cl))
        ar = np.maximum((v1 + a), (wr[IVX] + cr))
        bp = np.where(ar > 0.0, ar, 0.0)
        bm = np.where(al < 0.0, al, 0.0)
        vxl = wl[IVX] - bm
        vxr = wr[IVX] - bp
        fl = [wl[IDN] * vxl, wl[IDN] * wl[IVX] * vxl + pbl - _sq(bxi) + wl[IPR],
              wl[IDN] * wl[IVY] * vxl - bxi * wl[IBY], wl[IDN] * wl[IVZ] * vxl - bxi * wl[IBZ],
              el * vxl + wl[IVX] * (wl[IPR] + pbl - bxi * bxi) - bxi * (wl[IBY] * wl[IVY] + wl[IBZ] * wl[IVZ]),
              wl[IBY] * vxl - bxi * wl[IVY], wl[IBZ] * vxl - bxi * wl[IVZ]]
        fr = [wr[IDN] * vxr, wr[IDN] * wr[IVX] * vxr + pbr - _sq(bxi) + wr[IPR],
              wr[IDN] * wr[IVY] * vxr - bxi * wr[IBY], wr[IDN] * wr[IVZ] * vxr - bxi * wr[IBZ],
              er * vxr + wr[IVX] * (wr[IPR] + pbr - bxi * bxi) - bxi * (wr[IBY] * wr[IVY] + wr[IBZ] * wr[IVZ]),
              wr[IBY] * vxr - bxi * wr[IVY], wr[IBZ] * vxr - bxi * wr[IVZ]]
        tmp = np.where(bp != bm, 0.5 * (bp + bm) / (bp - bm), 0.0)
        return [0.5 * (fl[n] + fr[n]) + (fl[n] - fr[n]) * tmp for n in range(7)]


def roe_mhd(wl, wr, bxi, gamma):   # hydro/rsolvers/mhd/roe_mhd.cpp:60-215 + RoeFlux :245-470 (adiabatic)
    gm1 = gamma - 1.0
    with np.errstate(all="ignore"):
        sqrtdl, sqrtdr = np.sqrt(wl[IDN]), np.sqrt(wr[IDN])
        isdlpdr = 1.0 / (sqrtdl + sqrtdr)
        d = sqrtdl * sqrtdr
        v1, v2, v3 = [(sqrtdl * wl[v] + sqrtdr * wr[v]) * isdlpdr for v in (IVX, IVY, IVZ)]
        b2 = (sqrtdr * wl[IBY] + sqrtdl * wr[IBY]) * isdlpdr
        b3 = (sqrtdr * wl[IBZ] + sqrtdl * wr[IBZ]) * isdlpdr
        x = 0.5 * (_sq(wl[IBY] - wr[IBY]) + _sq(wl[IBZ] - wr[IBZ])) / (_sq(sqrtdl + sqrtdr))
        y = 0.5 * (wl[IDN] + wr[IDN]) / d
        pbl = 0.5 * (bxi * bxi + _sq(wl[IBY]) + _sq(wl[IBZ]))
        pbr = 0.5 * (bxi * bxi + _sq(wr[IBY]) + _sq(wr[IBZ]))
        fl, fr, du = _mhd_lr(wl, wr, bxi, gamma)
        el = wl[IPR] / gm1 + 0.5 * wl[IDN] * (_sq(wl[IVX]) + _sq(wl[IVY]) + _sq(wl[IVZ])) + pbl
        er = wr[IPR] / gm1 + 0.5 * wr[IDN] * (_sq(wr[IVX]) + _sq(wr[IVY]) + _sq(wr[IVZ])) + pbr
        h = ((el + wl[IPR] + pbl) / sqrtdl + (er + wr[IPR] + pbr) / sqrtdr) * isdlpdr
        flx = [0.5 * (fl[n] + fr[n]) for n in range(7)]
        # RoeFlux
        b1 = bxi
        di = 1.0 / d
        btsq = b2 * b2 + b3 * b3
        vaxsq = b1 * b1 * di
        vsq = v1 * v1 + v2 * v2 + v3 * v3
        hp = h - (vaxsq + btsq * di)
        bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq
        twid_csq = np.maximum((gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x), TINY_NUMBER)
        ct2 = bt_starsq * di
        tsum = vaxsq + ct2 + twid_csq
        tdif = vaxsq + ct2 - twid_csq
        cf2_cs2 = np.sqrt(tdif * tdif + 4.0 * twid_csq * ct2)
        cfsq = 0.5 * (tsum + cf2_cs2)
        cf = np.sqrt(cfsq)
        cssq = twid_csq * vaxsq / cfsq
        cs = np.sqrt(cssq)
        bt = np.sqrt(btsq)
        bt_star = np.sqrt(bt_starsq)
        bet2 = np.where(bt != 0.0, b2 / bt, 0.0)
        bet3 = np.where(bt != 0.0, b3 / bt, 0.0)
        bet2_star = bet2 / np.sqrt(gm1 - (gm1 - 1.0) * y)
        bet3_star = bet3 / np.sqrt(gm1 - (gm1 - 1.0) * y)
        bet_starsq = bet2_star * bet2_star + bet3_star * bet3_star
        vbet = v2 * bet2_star + v3 * bet3_star
        q2_star = np.where(bet_starsq != 0.0, bet2_star / bet_starsq, 0.0)
        q3_star = np.where(bet_starsq != 0.0, bet3_star / bet_starsq, 0.0)
        c1 = (cfsq - cssq) <= 0.0
        c2 = (twid_csq - cssq) <= 0.0
        c3 = (cfsq - twid_csq) <= 0.0
        af_gen = np.sqrt((twid_csq - cssq) / (cfsq - cssq))
        as_gen = np.sqrt((cfsq - twid_csq) / (cfsq - cssq))
        alpha_f = np.where(c1, 1.0, np.where(c2, 0.0, np.where(c3, 1.0, af_gen)))
        alpha_s = np.where(c1, 0.0, np.where(c2, 1.0, np.where(c3, 0.0, as_gen)))
        sqrtd = np.sqrt(d)
        isqrtd = 1.0 / sqrtd
        sg = sign(b1)
        twid_c = np.sqrt(twid_csq)
        qf = cf * alpha_f * sg
        qs = cs * alpha_s * sg
        af_prime = twid_c * alpha_f * isqrtd
        as_prime = twid_c * alpha_s * isqrtd
        afpbb = af_prime * bt_star * bet_starsq
        aspbb = as_prime * bt_star * bet_starsq
        vqstr = (v2 * q2_star + v3 * q3_star)
        vax = np.sqrt(vaxsq)
        norm = 0.5 / twid_csq
        cff = norm * alpha_f * cf
        css = norm * alpha_s * cs
        qf_hat = qf * norm
        qs_hat = qs * norm
        af = norm * af_prime * d
        as_ = norm * as_prime * d
        afpb = norm * af_prime * bt_star
        aspb = norm * as_prime * bt_star
        ev = [v1 - cf, v1 - vax, v1 - cs, v1, v1 + cs, v1 + vax, v1 + cf]
        alpha_f_bar = alpha_f * gm1 * norm
        alpha_s_bar = alpha_s * gm1 * norm
        gm1a = gm1 / twid_csq
        a = [None] * 7
        t = du[0] * (alpha_f_bar * (vsq - hp) + cff * (cf + v1) - qs_hat * vqstr - aspb)
        t = t - du[1] * (alpha_f_bar * v1 + cff)
        t = t - du[2] * (alpha_f_bar * v2 - qs_hat * q2_star)
        t = t - du[3] * (alpha_f_bar * v3 - qs_hat * q3_star)
        t = t + du[4] * alpha_f_bar
        t = t + du[5] * (as_ * q2_star - alpha_f_bar * b2)
        a[0] = t + du[6] * (as_ * q3_star - alpha_f_bar * b3)
        t = du[0] * (v2 * bet3 - v3 * bet2)
        t = t - du[2] * bet3
        t = t + du[3] * bet2
        t = t - du[5] * sqrtd * bet3 * sg
        t = t + du[6] * sqrtd * bet2 * sg
        a[1] = t * 0.5
        t = du[0] * (alpha_s_bar * (vsq - hp) + css * (cs + v1) + qf_hat * vqstr + afpb)
        t = t - du[1] * (alpha_s_bar * v1 + css)
        t = t - du[2] * (alpha_s_bar * v2 + qf_hat * q2_star)
        t = t - du[3] * (alpha_s_bar * v3 + qf_hat * q3_star)
        t = t + du[4] * alpha_s_bar
        t = t - du[5] * (af * q2_star + alpha_s_bar * b2)
        a[2] = t - du[6] * (af * q3_star + alpha_s_bar * b3)
        t = du[0] * (1.0 - gm1a * (0.5 * vsq - (gm1 - 1.0) * x / gm1))
        t = t + du[1] * gm1a * v1
        t = t + du[2] * gm1a * v2
        t = t + du[3] * gm1a * v3
        t = t - du[4] * gm1a
        t = t + du[5] * gm1a * b2
        a[3] = t + du[6] * gm1a * b3
        t = du[0] * (alpha_s_bar * (vsq - hp) + css * (cs - v1) - qf_hat * vqstr + afpb)
        t = t - du[1] * (alpha_s_bar * v1 - css)
        t = t - du[2] * (alpha_s_bar * v2 - qf_hat * q2_star)
        t = t - du[3] * (alpha_s_bar * v3 - qf_hat * q3_star)
        t = t + du[4] * alpha_s_bar
        t = t - du[5] * (af * q2_star + alpha_s_bar * b2)
        a[4] = t - du[6] * (af * q3_star + alpha_s_bar * b3)
        t = du[0] * (v3 * bet2 - v2 * bet3)
        t = t + du[2] * bet3
        t = t - du[3] * bet2
        t = t - du[5] * sqrtd * bet3 * sg
        t = t + du[6] * sqrtd * bet2 * sg
        a[5] = t * 0.5
        t = du[0] * (alpha_f_bar * (vsq - hp) + cff * (cf - v1) + qs_hat * vqstr - aspb)
        t = t - du[1] * (alpha_f_bar * v1 - cff)
        t = t - du[2] * (alpha_f_bar * v2 + qs_hat * q2_star)
        t = t - du[3] * (alpha_f_bar * v3 + qs_hat * q3_star)
        t = t + du[4] * alpha_f_bar
        t = t + du[5] * (as_ * q2_star - alpha_f_bar * b2)
        a[6] = t + du[6] * (as_ * q3_star - alpha_f_bar * b3)
        co = [-0.5 * np.abs(ev[n]) * a[n] for n in range(7)]
        dens = wl[IDN] + a[0] * alpha_f
        llf = dens < 0.0
        dens = dens + a[2] * alpha_s
        llf = llf | (dens < 0.0)
        dens = dens + a[3]
        llf = llf | (dens < 0.0)
        dens = dens + a[4] * alpha_s
        llf = llf | (dens < 0.0)
        f0 = flx[0] + co[0] * alpha_f
        f0 = f0 + co[2] * alpha_s
        f0 = f0 + co[3]
        f0 = f0 + co[4] * alpha_s
        f0 = f0 + co[6] * alpha_f
        f1 = flx[1] + co[0] * (alpha_f * (v1 - cf))
        f1 = f1 + co[2] * (alpha_s * (v1 - cs))
        f1 = f1 + co[3] * v1
        f1 = f1 + co[4] * (alpha_s * (v1 + cs))
        f1 = f1 + co[6] * (alpha_f * (v1 + cf))
        f2 = flx[2] + co[0] * (alpha_f * v2 + qs * bet2_star)
        f2 = f2 - co[1] * bet3
        f2 = f2 + co[2] * (alpha_s * v2 - qf * bet2_star)
        f2 = f2 + co[3] * v2
        f2 = f2 + co[4] * (alpha_s * v2 + qf * bet2_star)
        f2 = f2 + co[5] * bet3
        f2 = f2 + co[6] * (alpha_f * v2 - qs * bet2_star)
        f3 = flx[3] + co[0] * (alpha_f * v3 + qs * bet3_star)
        f3 = f3 + co[1] * bet2
        f3 = f3 + co[2] * (alpha_s * v3 - qf * bet3_star)
        f3 = f3 + co[3] * v3
        f3 = f3 + co[4] * (alpha_s * v3 + qf * bet3_star)
        f3 = f3 - co[5] * bet2
        f3 = f3 + co[6] * (alpha_f * v3 - qs * bet3_star)
        f4 = flx[4] + co[0] * (alpha_f * (hp - v1 * cf) + qs * vbet + aspbb)
        f4 = f4 - co[1] * (v2 * bet3 - v3 * bet2)
        f4 = f4 + co[2] * (alpha_s * (hp - v1 * cs) - qf * vbet - afpbb)
        f4 = f4 + co[3] * (0.5 * vsq + (gm1 - 1.0) * x / gm1)
        f4 = f4 + co[4] * (alpha_s * (hp + v1 * cs) + qf * vbet - afpbb)
        f4 = f4 + co[5] * (v1 * bet3 - v3 * bet2)
        f4 = f4 + co[6] * (alpha_f * (hp + v1 * cf) - qs * vbet + aspbb)
        f5 = flx[5] + co[0] * as_prime * bet2_star
        f5 = f5 - co[1] * bet3 * sg * isqrtd
        f5 = f5 - co[2] * af_prime * bet2_star
        f5 = f5 - co[4] * af_prime * bet2_star
        f5 = f5 - co[5] * bet3 * sg * isqrtd
        f5 = f5 + co[6] * as_prime * bet2_star
        f6 = flx[6] + co[0] * as_prime * bet3_star
        f6 = f6 + co[1] * bet2 * sg * isqrtd
        f6 = f6 - co[2] * af_prime * bet3_star
        f6 = f6 - co[4] * af_prime * bet3_star
        f6 = f6 + co[5] * bet2 * sg * isqrtd
        f6 = f6 + co[6] * as_prime * bet3_star
        out = [f0, f1, f2, f3, f4, f5, f6]
        out = [np.where(ev[0] >= 0.0, fl[n], out[n]) for n in range(7)]
        out = [np.where(ev[6] <= 0.0, fr[n], out[n]) for n in range(7)]
        cfl_ = fast_speed(gamma, wl[IDN], wl[IPR], wl[IBY], wl[IBZ], bxi)
        cfr_ = fast_speed(gamma, wr[IDN], wr[IPR], wr[IBY], wr[IBZ], bxi)
        al = 0.5 * np.maximum((np.abs(wl[IVX]) + cfl_), (np.abs(wr[IVX]) + cfr_))
        # the LLF fallback leaves the induction components at their Roe values (roe_mhd.cpp:193-205)
        return [np.where(llf, 0.5 * (fl[n] + fr[n]) - al * du[n], out[n]) if n < 5 else out[n] for n in range(7)]


HYDRO_SOLVERS = {"hllc": None, "hlle": hlle_hydro, "llf": llf_hydro, "roe": roe_hydro}
MHD_SOLVERS = {"hlld": None, "hlle": hlle_mhd, "llf": llf_mhd, "roe": roe_mhd}


def hlld(wl, wr, bxi, gamma):   # hlld.cpp:56-365 ; returns 7 flux components
    SMALL = 1.0e-4
    igm1 = 1.0 / (gamma - 1.0)
    with np.errstate(all="ignore"):
        bxsq = bxi * bxi
        pbl = 0.5 * (bxsq + (wl[IBY] * wl[IBY] + wl[IBZ] * wl[IBZ]))
        pbr = 0.5 * (bxsq + (wr[IBY] * wr[IBY] + wr[IBZ] * wr[IBZ]))
        kel = 0.5 * wl[IDN] * (wl[IVX] * wl[IVX] + (wl[IVY] * wl[IVY] + wl[IVZ] * wl[IVZ]))
        ker = 0.5 * wr[IDN] * (wr[IVX] * wr[IVX] + (wr[IVY] * wr[IVY] + wr[IVZ] * wr[IVZ]))
        # conserved L/R: d, mx, my, mz, e, by, bz
        ul = [wl[IDN], wl[IVX] * wl[IDN], wl[IVY] * wl[IDN], wl[IVZ] * wl[IDN],
              wl[IPR] * igm1 + kel + pbl, wl[IBY], wl[IBZ]]
        ur = [wr[IDN], wr[IVX] * wr[IDN], wr[IVY] * wr[IDN], wr[IVZ] * wr[IDN],
              wr[IPR] * igm1 + ker + pbr, wr[IBY], wr[IBZ]]
        cfl = fast_speed(gamma, wl[IDN], wl[IPR], wl[IBY], wl[IBZ], bxi)
        cfr = fast_speed(gamma, wr[IDN], wr[IPR], wr[IBY], wr[IBZ], bxi)
        a0, b0 = wl[IVX] - cfl, wr[IVX] - cfr
        s0 = np.where(b0 < a0, b0, a0)                 # std::min
        a4, b4 = wl[IVX] + cfl, wr[IVX] + cfr
        s4 = np.where(a4 < b4, b4, a4)                 # std::max
        ptl = wl[IPR] + pbl
        ptr = wr[IPR] + pbr
        fl = [ul[1], ul[1] * wl[IVX] + ptl - bxsq, ul[2] * wl[IVX] - bxi * ul[5],
              ul[3] * wl[IVX] - bxi * ul[6],
              wl[IVX] * (ul[4] + ptl - bxsq) - bxi * (wl[IVY] * ul[5] + wl[IVZ] * ul[6]),
              ul[5] * wl[IVX] - bxi * wl[IVY], ul[6] * wl[IVX] - bxi * wl[IVZ]]
        fr = [ur[1], ur[1] * wr[IVX] + ptr - bxsq, ur[2] * wr[IVX] - bxi * ur[5],
              ur[3] * wr[IVX] - bxi * ur[6],
              wr[IVX] * (ur[4] + ptr - bxsq) - bxi * (wr[IVY] * ur[5] + wr[IVZ] * ur[6]),
              ur[5] * wr[IVX] - bxi * wr[IVY], ur[6] * wr[IVX] - bxi * wr[IVZ]]
        sdl = s0 - wl[IVX]
        sdr = s4 - wr[IVX]
        s2 = (sdr * ur[1] - sdl * ul[1] + (ptl - ptr)) / (sdr * ur[0] - sdl * ul[0])
        sdml = s0 - s2
        sdmr = s4 - s2
        sdml_inv = 1.0 / sdml
        sdmr_inv = 1.0 / sdmr
        ulst_d = ul[0] * sdl * sdml_inv
        urst_d = ur[0] * sdr * sdmr_inv
        ulst_d_inv = 1.0 / ulst_d
        urst_d_inv = 1.0 / urst_d
        sqrtdl = np.sqrt(ulst_d)
        sqrtdr = np.sqrt(urst_d)
        s1 = s2 - np.abs(bxi) / sqrtdl
        s3 = s2 + np.abs(bxi) / sqrtdr
        ptstl = ptl + ul[0] * sdl * (s2 - wl[IVX])
        ptstr = ptr + ur[0] * sdr * (s2 - wr[IVX])
        ptst = 0.5 * (ptstr + ptstl)

        def star(u, w, sd, sdm, sdm_inv, ust_d, ust_d_inv, pt):
            mx = ust_d * s2
            den = u[0] * sd * sdm - bxsq
            degen = np.abs(den) < SMALL * ptst
            tmp = bxi * (sd - sdm) / den
            my = np.where(degen, ust_d * w[IVY], ust_d * (w[IVY] - u[5] * tmp))
            mz = np.where(degen, ust_d * w[IVZ], ust_d * (w[IVZ] - u[6] * tmp))
            tmp2 = (u[0] * (sd * sd) - bxsq) / den
            by = np.where(degen, u[5], u[5] * tmp2)
            bz = np.where(degen, u[6], u[6] * tmp2)
            vbst = (mx * bxi + (my * by + mz * bz)) * ust_d_inv
            e = (sd * u[4] - pt * w[IVX] + ptst * s2 +
                 bxi * (w[IVX] * bxi + (w[IVY] * u[5] + w[IVZ] * u[6]) - vbst)) * sdm_inv
            return [ust_d, mx, my, mz, e, by, bz], vbst

        ulst, vbstl = star(ul, wl, sdl, sdml, sdml_inv, ulst_d, ulst_d_inv, ptl)
        urst, vbstr = star(ur, wr, sdr, sdmr, sdmr_inv, urst_d, urst_d_inv, ptr)
        invsumd = 1.0 / (sqrtdl + sqrtdr)
        bxsig = np.where(bxi > 0.0, 1.0, -1.0)
        tmp = invsumd * (sqrtdl * (ulst[2] * ulst_d_inv) + sqrtdr * (urst[2] * urst_d_inv) +
                         bxsig * (urst[5] - ulst[5]))
        uld_my, urd_my = ulst[0] * tmp, urst[0] * tmp
        tmp = invsumd * (sqrtdl * (ulst[3] * ulst_d_inv) + sqrtdr * (urst[3] * urst_d_inv) +
                         bxsig * (urst[6] - ulst[6]))
        uld_mz, urd_mz = ulst[0] * tmp, urst[0] * tmp
        d_by = invsumd * (sqrtdl * urst[5] + sqrtdr * ulst[5] +
                          bxsig * sqrtdl * sqrtdr * ((urst[2] * urst_d_inv) - (ulst[2] * ulst_d_inv)))
        d_bz = invsumd * (sqrtdl * urst[6] + sqrtdr * ulst[6] +
                          bxsig * sqrtdl * sqrtdr * ((urst[3] * urst_d_inv) - (ulst[3] * ulst_d_inv)))
        tmp = s2 * bxi + (uld_my * d_by + uld_mz * d_bz) / ulst[0]
        uld_e = ulst[4] - sqrtdl * bxsig * (vbstl - tmp)
        urd_e = urst[4] + sqrtdr * bxsig * (vbstr - tmp)
        uldst = [ulst[0], ulst[1], uld_my, uld_mz, uld_e, d_by, d_bz]
        urdst = [urst[0], urst[1], urd_my, urd_mz, urd_e, d_by, d_bz]
        jl2 = [s1 * (uldst[n] - ulst[n]) for n in range(7)]
        jl1 = [s0 * (ulst[n] - ul[n]) for n in range(7)]
        jr2 = [s3 * (urdst[n] - urst[n]) for n in range(7)]
        jr1 = [s4 * (urst[n] - ur[n]) for n in range(7)]
        f = []
        for n in range(7):
            val = fr[n] + jr1[n] + jr2[n]                             # Fr**
            val = np.where(s2 >= 0.0, fl[n] + jl1[n] + jl2[n], val)   # Fl**
            val = np.where(s3 <= 0.0, fr[n] + jr1[n], val)            # Fr*
            val = np.where(s1 >= 0.0, fl[n] + jl1[n], val)            # Fl*
            val = np.where(s4 <= 0.0, fr[n], val)
            val = np.where(s0 >= 0.0, fl[n], val)
            f.append(val)
    return f


def ct_weight(dflx, rhol, rhor, dx, dt):   # hydro.cpp:154-158
    v_over_c = (1024.0) * dt * dflx / (dx * (rhol + rhor))
    tmp_min = np.where(v_over_c < 0.5, v_over_c, 0.5)
    return 0.5 + np.where(-0.5 < tmp_min, tmp_min, -0.5)


class State:
    """u: global interior (5,n3,n2,n1); b: global face arrays with the +1 extent of their own
    direction (the faces on the periodic seam are distinct arrays entries, as in the reference
    where each MeshBlock keeps its own copy of a shared face)."""

    def __init__(self, mesh, gamma, u, b=None, s=None, flux=None):
        self.mesh, self.gamma = mesh, gamma
        self.flux = flux   # None = hllc / hlld; "hlle", "llf", "roe" (hydro)
        self.mhd = b is not None
        self.u = mesh.fill_ghosts_cc(mesh.pad_cc(u))
        self.b = [mesh.fill_ghosts_fc(mesh.pad_fc(b[d], d), d) for d in range(3)] if self.mhd else None
        # passive scalars: conserved s = r*rho (NSCALARS, ...), scalars.hpp
        self.s = mesh.fill_ghosts_cc(mesh.pad_cc(s)) if s is not None else None
        self.prim()

    def prim(self):
        self.bcc = cell_centered_b(*self.b) if self.mhd else None
        with np.errstate(all="ignore"):
            self.w = cons2prim(self.u, self.gamma, self.bcc)
            if self.s is not None:   # PassiveScalarConservedToPrimitive, eos_scalars.cpp:31-60
                d = self.u[IDN]
                self.s = np.where(self.s < FLOOR * d, FLOOR * d, self.s)
                self.r = self.s / d
            apply_physical_bcs(self)

    def interior_s(self):
        return self.mesh.unpad(self.s)

    def interior_u(self):
        return self.mesh.unpad(self.u)

    def interior_b(self, d):
        return self.mesh.unpad(self.b[d], d)


def fluxes(st, order, dt_mesh):
    """CalculateFluxes: returns flux[d] (5 arrays, variable-indexed) and, for MHD, face EMFs and
    CT weights, all at face n of direction d (between cells n-1 and n)."""
    mesh, g = st.mesh, st.gamma
    out = []
    for d in range(3):
        ivx, ivy, ivz = IVX + d, IVX + (d + 1) % 3, IVX + (d + 2) % 3
        q = [st.w[IDN], st.w[ivx], st.w[ivy], st.w[ivz], st.w[IPR]]
        if st.mhd:
            q += [st.bcc[(d + 1) % 3], st.bcc[(d + 2) % 3]]
        wl, wr = reconstruct(q, d, order, mesh)
        if order == 3:   # ApplyPrimitiveFloors on both states, ppm.cpp:326-332
            for ws in (wl, wr):
                ws[0] = np.where(ws[0] > FLOOR, ws[0], FLOOR)
                ws[4] = np.where(ws[4] > FLOOR, ws[4], FLOOR)
        if st.mhd:
            f = (MHD_SOLVERS.get(st.flux) or hlld)(wl, wr, st.b[d], g)
            flx = [None] * 5
            flx[IDN], flx[ivx], flx[ivy], flx[ivz], flx[IEN] = f[0], f[1], f[2], f[3], f[4]
            dxw = mesh.along(mesh.dxf[d], d)
            wct = ct_weight(f[IDN], wl[IDN], wr[IDN], dxw, dt_mesh)
            out.append((flx, -f[IBY], f[IBZ], wct))
        else:
            f = (HYDRO_SOLVERS.get(st.flux) or hllc)(wl, wr, g)
            flx = [None] * 5
            flx[IDN], flx[ivx], flx[ivy], flx[ivz], flx[IEN] = f[0], f[1], f[2], f[3], f[4]
            out.append((flx, None, None, None))
    return out


def scalar_fluxes(st, fx, order):
    """PassiveScalars::CalculateFluxes + ComputeUpwindFlux (calculate_scalar_fluxes.cpp:41-417):
    reconstruct r = s/rho like a hydro variable, concentration floor on both states for xorder=3,
    flux = F(rho) * (F(rho) >= 0 ? r_L : r_R) with the hydro mass flux of the same face."""
    out = []
    for d in range(3):
        rl, rr = reconstruct(list(st.r), d, order, st.mesh)
        if order == 3:
            rl = [np.where(x > FLOOR, x, FLOOR) for x in rl]
            rr = [np.where(x > FLOOR, x, FLOOR) for x in rr]
        mf = fx[d][0][IDN]
        out.append([np.where(mf >= 0.0, mf * a, mf * b) for a, b in zip(rl, rr)])
    return out


def corner_e(st, fx):
    """ComputeCornerE (3-D) + same-level EMF correction for a periodic uniform mesh.  On the
    global mesh every edge is interior, so the correction (own + bit-equal neighbour copies,
    then *0.5 or *0.25) is the identity except for rounding of ((a+a)+a)+a; the reference applies
    it only on MeshBlock boundaries, where it is exact as well -- see apply_emf_correction."""
    w, bcc = st.w, st.bcc
    cc1 = w[IVZ] * bcc[1] - w[IVY] * bcc[2]
    cc2 = w[IVX] * bcc[2] - w[IVZ] * bcc[0]
    cc3 = w[IVY] * bcc[0] - w[IVX] * bcc[1]
    (_, e3_x1f, e2_x1f, w_x1f), (_, e1_x2f, e3_x2f, w_x2f), (_, e2_x3f, e1_x3f, w_x3f) = fx
    S = sh

    def K(a, n=-1):
        return S(a, 2, n)

    def J(a, n=-1):
        return S(a, 1, n)

    def I(a, n=-1):
        return S(a, 0, n)
    de1_l3 = (1.0 - K(w_x2f)) * (e1_x3f - K(cc1)) + (K(w_x2f)) * (J(e1_x3f) - K(J(cc1)))
    de1_r3 = (1.0 - w_x2f) * (e1_x3f - cc1) + (w_x2f) * (J(e1_x3f) - J(cc1))
    de1_l2 = (1.0 - J(w_x3f)) * (e1_x2f - J(cc1)) + (J(w_x3f)) * (K(e1_x2f) - K(J(cc1)))
    de1_r2 = (1.0 - w_x3f) * (e1_x2f - cc1) + (w_x3f) * (K(e1_x2f) - K(cc1))
    e1 = 0.25 * (de1_l3 + de1_r3 + de1_l2 + de1_r2 + K(e1_x2f) + e1_x2f + J(e1_x3f) + e1_x3f)
    de2_l3 = (1.0 - K(w_x1f)) * (e2_x3f - K(cc2)) + (K(w_x1f)) * (I(e2_x3f) - K(I(cc2)))
    de2_r3 = (1.0 - w_x1f) * (e2_x3f - cc2) + (w_x1f) * (I(e2_x3f) - I(cc2))
    de2_l1 = (1.0 - I(w_x3f)) * (e2_x1f - I(cc2)) + (I(w_x3f)) * (K(e2_x1f) - K(I(cc2)))
    de2_r1 = (1.0 - w_x3f) * (e2_x1f - cc2) + (w_x3f) * (K(e2_x1f) - K(cc2))
    e2 = 0.25 * (de2_l3 + de2_r3 + de2_l1 + de2_r1 + I(e2_x3f) + e2_x3f + K(e2_x1f) + e2_x1f)
    de3_l2 = (1.0 - J(w_x1f)) * (e3_x2f - J(cc3)) + (J(w_x1f)) * (I(e3_x2f) - J(I(cc3)))
    de3_r2 = (1.0 - w_x1f) * (e3_x2f - cc3) + (w_x1f) * (I(e3_x2f) - I(cc3))
    de3_l1 = (1.0 - I(w_x2f)) * (e3_x1f - I(cc3)) + (I(w_x2f)) * (J(e3_x1f) - J(I(cc3)))
    de3_r1 = (1.0 - w_x2f) * (e3_x1f - cc3) + (w_x2f) * (J(e3_x1f) - J(cc3))
    e3 = 0.25 * (de3_l1 + de3_r1 + de3_l2 + de3_r2 + I(e3_x2f) + e3_x2f + J(e3_x1f) + e3_x1f)
    return [e1, e2, e3]


def apply_emf_correction(e, mesh, block_nx):
    """SetFluxBoundarySameLevel + AverageFluxBoundary on MeshBlock boundaries
    (flux_correction_fc.cpp:689-760,1355-1541): every block adds its neighbours' copies of a
    shared edge EMF to its own (face neighbours in x1,x2,x3 order, then the edge neighbour),
    then scales faces by 0.5 and block edges by 0.25.  Inside the mesh all copies are the same
    array entry; across the periodic seam the partner copy is the entry at the opposite seam
    index (it was computed from that side's own copy of the seam face field, so it may differ in
    the last bits)."""
    ng = mesh.ng

    def partner(a, d):
        i, n = mesh.idx(d), mesh.nx[d]
        return np.where(i == ng, np.roll(a, -n, axis=AX[d]),
                        np.where(i == ng + n, np.roll(a, n, axis=AX[d]), a))
    on = []
    for d in range(3):
        i = mesh.idx(d) - ng
        shared = (i >= 0) & (i <= mesh.nx[d]) & (i % block_nx[d] == 0)
        if mesh.bc[d][0] is not None:   # the mesh edge has no neighbour: nothing to add or average
            shared = shared & (i != 0) & (i != mesh.nx[d])
        on.append(shared)
    out = []
    for comp, (da, db) in enumerate(((1, 2), (0, 2), (0, 1))):
        a = e[comp]
        pa, pb = on[da], on[db]
        fa = partner(a, da)
        fb = partner(a, db)
        fe = partner(fa, db)
        v_a = (a + fa) * 0.5
        v_b = (a + fb) * 0.5
        v_edge = (((a + fa) + fb) + fe) * 0.25
        out.append(np.where(pa & pb, v_edge, np.where(pa, v_a, np.where(pb, v_b, a))))
    return out


def apply_physical_bcs(st):
    """BoundaryValues::ApplyPhysicalBoundaries (bvals/bvals.cpp:453-660) with the built-in
    reflecting / outflow functions (bvals/cc/hydro/reflect_hydro.cpp, bvals/cc/outflow_cc.cpp,
    bvals/cc/reflect_cc.cpp for r, bvals/fc/reflect_fc.cpp, outflow_fc.cpp) on w, r and b; then
    bcc, PrimitiveToConserved (adiabatic_*.cpp) and s = r*rho on the same ghost slab.  Faces in
    the order x1, x2, x3; a transverse direction includes its ghost zones once they are valid
    (periodic from the exchange, or physical and already processed: bvals.cpp:457-470,560)."""
    m = st.mesh
    ng = m.ng
    if all(m.bc[d][0] is None for d in range(3)):
        return
    igm1 = 1.0 / (st.gamma - 1.0)

    def gather(a, d, src):   # a[..., index along d] <- a[..., src[index]]
        return np.take(a, src, axis=AX[d])

    for d in range(3):
        n = m.nx[d]
        idx1 = np.arange(m.nc[d] + 1)
        for side in (0, 1):
            flag = m.bc[d][side]
            if flag is None:
                continue
            refl = flag == "reflecting"
            i = m.idx(d)
            ghost_c = (i < ng) if side == 0 else ((i > ng + n - 1) & (i <= ng + n - 1 + ng))
            ghost_f = (i < ng) if side == 0 else ((i > ng + n) & (i <= ng + n + ng))
            if side == 0:
                src_c = np.where(idx1 < ng, (2 * ng - 1 - idx1) if refl else ng, idx1)
                src_f = np.where(idx1 < ng, (2 * ng - idx1) if refl else ng, idx1)
            else:
                src_c = np.where(idx1 > ng + n - 1, (2 * (ng + n) - 1 - idx1) if refl else ng + n - 1, idx1)
                src_f = np.where(idx1 > ng + n, (2 * (ng + n) - idx1) if refl else ng + n, idx1)
            src_c = np.clip(src_c, 0, m.nc[d])
            src_f = np.clip(src_f, 0, m.nc[d])
            # transverse ranges of the cell-centred pass
            tr = []
            for t in range(3):
                if t == d:
                    continue
                it = m.idx(t)
                lo = ng - (ng if (t < d or m.bc[t][0] is None) else 0)
                hi = ng + m.nx[t] - 1 + (ng if (t < d or m.bc[t][1] is None) else 0)
                tr.append((t, it, lo, hi))
            cmask = ghost_c
            for t, it, lo, hi in tr:
                cmask = cmask & (it >= lo) & (it <= hi)
            for v in range(5):
                val = gather(st.w[v], d, src_c)
                if refl and v == IVX + d:
                    val = -val
                st.w[v] = np.where(cmask, val, st.w[v])
            if st.s is not None:
                for q in range(st.r.shape[0]):
                    st.r[q] = np.where(cmask, gather(st.r[q], d, src_c), st.r[q])
            if st.mhd:
                for comp in range(3):
                    fm = ghost_f if comp == d else ghost_c
                    for t, it, lo, hi in tr:
                        fm = fm & (it >= lo) & (it <= hi + (1 if t == comp else 0))
                    val = gather(st.b[comp], d, src_f if comp == d else src_c)
                    if refl and comp == d:
                        val = -val
                    st.b[comp] = np.where(fm, val, st.b[comp])
                bcc = cell_centered_b(*st.b)
                st.bcc = tuple(np.where(cmask, bcc[c], st.bcc[c]) for c in range(3))
            w = st.w
            e = w[IPR] * igm1 + 0.5 * w[IDN] * (w[IVX] * w[IVX] + w[IVY] * w[IVY] + w[IVZ] * w[IVZ])
            if st.mhd:
                e = w[IPR] * igm1 + 0.5 * (w[IDN] * (w[IVX] * w[IVX] + w[IVY] * w[IVY] + w[IVZ] * w[IVZ])
                                           + (st.bcc[0] * st.bcc[0] + st.bcc[1] * st.bcc[1] + st.bcc[2] * st.bcc[2]))
            unew = [w[IDN], w[IVX] * w[IDN], w[IVY] * w[IDN], w[IVZ] * w[IDN], e]
            for v in range(5):
                st.u[v] = np.where(cmask, unew[v], st.u[v])
            if st.s is not None:
                st.s = np.where(cmask, st.r * w[IDN], st.s)


def weighted_ave(out, in1, in2, w0, w1, w2):   # weighted_ave.cpp:35-230 (wght[3]=wght[4]=0)
    if w0 == 1.0:
        if w2 != 0.0:
            return out + (w1 * in1 + w2 * in2)
        if w1 != 0.0:
            return out + w1 * in1
        return out
    if w0 == 0.0:
        if w2 != 0.0:
            return w1 * in1 + w2 * in2
        if w1 != 0.0:
            return w1 * in1
        return out
    if w2 != 0.0:
        return w0 * out + w1 * in1 + w2 * in2
    if w1 != 0.0:
        return w0 * out + w1 * in1
    return out * w0


def block_dt(st, cfl):   # new_blockdt.cpp:72-191
    w, m = st.w, st.mesh
    dx = [m.along(m.dxf[d], d) for d in range(3)]
    with np.errstate(all="ignore"):
        if st.mhd:
            bcc, b = st.bcc, st.b
            dts = []
            for d in range(3):
                bx = bcc[d] + np.abs(b[d] - bcc[d])
                cf = fast_speed(st.gamma, w[IDN], w[IPR], bcc[(d + 1) % 3], bcc[(d + 2) % 3], bx)
                dts.append(dx[d] / np.maximum(0.0, np.abs(w[IVX + d]) + cf))
        else:
            cs = np.sqrt(st.gamma * w[IPR] / w[IDN])
            dts = [dx[d] / np.maximum(0.0, np.abs(w[IVX + d]) + cs) for d in range(3)]
    act = m.active(0) & m.active(1) & m.active(2)
    return min(float(x[np.broadcast_to(act, x.shape)].min()) for x in dts) * cfl


class Integrator:
    def __init__(self, st, integrator="vl2", xorder=2, cfl=0.3, block_nx=None, grav_acc=None):
        self.st, self.name, self.xorder = st, integrator, xorder
        self.grav_acc = tuple(float(x) for x in (grav_acc or (0.0, 0.0, 0.0)))   # <hydro> grav_acc1/2/3
        self.stages = INTEGRATORS[integrator]
        self.cfl = min(cfl, CFL_LIMIT[integrator])
        self.block_nx = block_nx or st.mesh.nx
        self.time, self.ncycle = 0.0, 0
        self.dt = sys.float_info.max
        self.dt_history = []

    def new_dt(self, tlim):   # mesh.cpp:1236-1282
        dt = min(2.0 * self.dt, block_dt(self.st, self.cfl))
        if tlim is not None and self.time < tlim and (tlim - self.time) < dt:
            dt = tlim - self.time
        self.dt = dt

    def step(self, tlim=None):
        st, m = self.st, self.st.mesh
        dx = [m.along(m.dxf[d], d) for d in range(3)]
        act = [m.active(d) for d in range(3)]
        actf = [m.active(d, 1) for d in range(3)]
        cells = act[0] & act[1] & act[2]
        # WeightedAve / CT ranges of the face fields (weighted_ave.cpp:383-406, ct.cpp:31-116)
        faces = [actf[0] & act[1] & act[2], act[0] & actf[1] & act[2], act[0] & act[1] & actf[2]]
        u1 = np.zeros_like(st.u)
        b1 = [np.zeros_like(x) for x in st.b] if st.mhd else None
        s1 = np.zeros_like(st.s) if st.s is not None else None
        for stage, (delta, g1, g2, g3, beta) in enumerate(self.stages, 1):
            order = 1 if (self.name == "vl2" and stage == 1) else self.xorder
            with np.errstate(all="ignore"):
                fx = fluxes(st, order, self.dt)
                wght = beta * self.dt
                if st.mhd:
                    e = apply_emf_correction(corner_e(st, fx), m, self.block_nx)
                # IntegrateHydro (time_integrator.cpp:1797-1824); only active cells are touched
                u1 = np.where(cells, weighted_ave(u1, st.u, None, 1.0, delta, 0.0), u1)
                if g1 == 0.0 and g2 == 1.0 and g3 == 0.0:
                    st.u, u1 = u1, st.u
                else:
                    st.u = np.where(cells, weighted_ave(st.u, u1, None, g1, g2, g3), st.u)
                a1, a2, a3 = dx[1] * dx[2], dx[0] * dx[2], dx[0] * dx[1]
                vol = dx[0] * dx[1] * dx[2]
                unew = st.u.copy()
                for n in range(5):
                    dflx = (a1 * sh(fx[0][0][n], 0, 1) - a1 * fx[0][0][n])
                    dflx = dflx + (a2 * sh(fx[1][0][n], 1, 1) - a2 * fx[1][0][n])
                    dflx = dflx + (a3 * sh(fx[2][0][n], 2, 1) - a3 * fx[2][0][n])
                    unew[n] = np.where(cells, st.u[n] - wght * dflx / vol, st.u[n])
                st.u = unew
                # AddSourceTerms -> ConstantAcceleration (hydro/srcterms/constant_acc.cpp:25-76): one
                # direction after the other, rho and v from the stage-input primitives
                for d, gd in enumerate(self.grav_acc):
                    if gd != 0.0:
                        src = wght * st.w[0] * gd
                        st.u[1 + d] = np.where(cells, st.u[1 + d] + src, st.u[1 + d])
                        st.u[4] = np.where(cells, st.u[4] + src * st.w[1 + d], st.u[4])
                if st.s is not None:   # IntegrateScalars (time_integrator.cpp:2376-2403)
                    sfx = scalar_fluxes(st, fx, order)
                    s1 = np.where(cells, weighted_ave(s1, st.s, None, 1.0, delta, 0.0), s1)
                    if g1 == 0.0 and g2 == 1.0 and g3 == 0.0:
                        st.s, s1 = s1, st.s
                    else:
                        st.s = np.where(cells, weighted_ave(st.s, s1, None, g1, g2, g3), st.s)
                    snew = st.s.copy()
                    for n in range(st.s.shape[0]):
                        dflx = (a1 * sh(sfx[0][n], 0, 1) - a1 * sfx[0][n])
                        dflx = dflx + (a2 * sh(sfx[1][n], 1, 1) - a2 * sfx[1][n])
                        dflx = dflx + (a3 * sh(sfx[2][n], 2, 1) - a3 * sfx[2][n])
                        snew[n] = np.where(cells, st.s[n] - wght * dflx / vol, st.s[n])
                    st.s = m.fill_ghosts_cc(snew)
                if st.mhd:   # IntegrateField + CT (ct.cpp:31-116)
                    for c in range(3):
                        b1[c] = np.where(faces[c], weighted_ave(b1[c], st.b[c], None, 1.0, delta, 0.0), b1[c])
                    if g1 == 0.0 and g2 == 1.0 and g3 == 0.0:
                        st.b, b1 = b1, st.b
                    else:
                        st.b = [np.where(faces[c], weighted_ave(st.b[c], b1[c], None, g1, g2, g3), st.b[c])
                                for c in range(3)]
                    e1, e2, e3 = e
                    bx1 = st.b[0] - (wght / a1) * (dx[2] * sh(e3, 1, 1) - dx[2] * e3)
                    bx1 = bx1 + (wght / a1) * (dx[1] * sh(e2, 2, 1) - dx[1] * e2)
                    bx2 = st.b[1] + (wght / a2) * (dx[2] * sh(e3, 0, 1) - dx[2] * e3)
                    bx2 = bx2 - (wght / a2) * (dx[0] * sh(e1, 2, 1) - dx[0] * e1)
                    bx3 = st.b[2] - (wght / a3) * (dx[1] * sh(e2, 0, 1) - dx[1] * e2)
                    bx3 = bx3 + (wght / a3) * (dx[0] * sh(e1, 1, 1) - dx[0] * e1)
                    st.b = [np.where(faces[0], bx1, st.b[0]), np.where(faces[1], bx2, st.b[1]),
                            np.where(faces[2], bx3, st.b[2])]
                    st.b = [m.fill_ghosts_fc(st.b[d], d) for d in range(3)]
                st.u = m.fill_ghosts_cc(st.u)
            st.prim()
        self.ncycle += 1
        self.time += self.dt
        self.dt_history.append(self.dt)
        self.new_dt(tlim)

    def run(self, tlim):
        self.new_dt(tlim)
        while self.time < tlim:
            self.step(tlim)
