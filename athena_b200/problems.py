"""Synthetic initial conditions for bench.py / tests, evaluated per MeshBlock on the host and then
uploaded (the reference runs its problem generators on the host too).  The formulas follow the
reference pgens (file:line under /root/reference/src/pgen):
  blast.cpp:52-259 (spherical blast, Cartesian branch).
The C++ host (athena_b200/csrc/host/pgen_*.cpp) implements the same pgens behind
MeshBlock::ProblemGenerator; this module is the numpy form used by the Python driver."""
import numpy as np

PI = 3.1415926535897932


def blast_block(mesh, loc, mhd, gamma, pamb=0.1, prat=100.0, radius=0.1, damb=1.0, drat=1.0,
                ramp=0.0, b0=1.0, angle_deg=30.0, center=(0.0, 0.0, 0.0)):
    """Returns dict(u, b1, b2, b3) for one MeshBlock, AthenaArray layout with ghosts (zero)."""
    ng = mesh.ng
    nx1, nx2, nx3 = mesh.block_nx
    xs = [mesh.coords(loc, d) for d in range(3)]
    x1v, x2v, x3v = xs[0][1], xs[1][1], xs[2][1]
    nc1, nc2, nc3 = nx1 + 2 * ng, nx2 + 2 * ng, nx3 + 2 * ng
    rout = radius
    rin = rout - ramp
    gm1 = gamma - 1.0
    X = x1v[None, None, :] - center[0]
    Y = x2v[None, :, None] - center[1]
    Z = x3v[:, None, None] - center[2]
    rad = np.sqrt(X * X + Y * Y + Z * Z)
    den = np.full((nc3, nc2, nc1), damb)
    pres = np.full((nc3, nc2, nc1), pamb)
    inner = rad < rin
    den[inner] = drat * damb
    pres[inner] = prat * pamb
    if ramp > 0.0:
        mid = (rad < rout) & ~inner
        f = (rad[mid] - rin) / (rout - rin)
        den[mid] = np.exp((1.0 - f) * np.log(drat * damb) + f * np.log(damb))
        pres[mid] = np.exp((1.0 - f) * np.log(prat * pamb) + f * np.log(pamb))
    u = np.zeros((5, nc3, nc2, nc1))
    sl = (slice(ng, ng + nx3), slice(ng, ng + nx2), slice(ng, ng + nx1))
    u[0][sl] = den[sl]
    u[4][sl] = pres[sl] / gm1
    out = dict(u=u)
    if mhd:
        ang = (PI / 180.0) * angle_deg
        b1 = np.zeros((nc3, nc2, nc1 + 1))
        b2 = np.zeros((nc3, nc2 + 1, nc1))
        b3 = np.zeros((nc3 + 1, nc2, nc1))
        b1[ng:ng + nx3, ng:ng + nx2, ng:ng + nx1 + 1] = b0 * np.cos(ang)
        b2[ng:ng + nx3, ng:ng + nx2 + 1, ng:ng + nx1] = b0 * np.sin(ang)
        u[4][sl] += 0.5 * b0 * b0
        out.update(b1=b1, b2=b2, b3=b3)
    return out


def sound_wave_block(mesh, loc, mhd, gamma, amp=1.0e-6, b=(1.0, 1.4142135623730951, 0.5)):
    """A small-amplitude plane wave along x1 on a uniform background (the shape of pgen/linear_wave.cpp's
    wave_flag=0 set-up: rho0 = 1, p0 = 1/gamma, |delta| ~ amp; with MHD a uniform field b): throughput
    workload for the BASELINE configs[0] / configs[1] shapes, not the eigenvector of the parity tests
    (those run through the C++ host's linear_wave problem generator)."""
    ng = mesh.ng
    nx1, nx2, nx3 = mesh.block_nx
    x1v = mesh.coords(loc, 0)[1]
    nc1, nc2, nc3 = nx1 + 2 * ng, nx2 + 2 * ng, nx3 + 2 * ng
    lx = mesh.xmax[0] - mesh.xmin[0]
    sn = np.sin(2.0 * PI * (x1v - mesh.xmin[0]) / lx)[None, None, :] * np.ones((nc3, nc2, 1))
    d = 1.0 + amp * sn
    p = 1.0 / gamma + amp * sn
    vx = amp * sn
    u = np.zeros((5, nc3, nc2, nc1))
    sl = (slice(ng, ng + nx3), slice(ng, ng + nx2), slice(ng, ng + nx1))
    u[0][sl] = d[sl]
    u[1][sl] = (d * vx)[sl]
    u[4][sl] = (p / (gamma - 1.0) + 0.5 * d * vx * vx)[sl]
    out = dict(u=u)
    if mhd:
        b1 = np.zeros((nc3, nc2, nc1 + 1))
        b2 = np.zeros((nc3, nc2 + 1, nc1))
        b3 = np.zeros((nc3 + 1, nc2, nc1))
        b1[ng:ng + nx3, ng:ng + nx2, ng:ng + nx1 + 1] = b[0]
        b2[ng:ng + nx3, ng:ng + nx2 + 1, ng:ng + nx1] = b[1]
        b3[ng:ng + nx3 + 1, ng:ng + nx2, ng:ng + nx1] = b[2]
        u[4][sl] += 0.5 * (b[0] ** 2 + b[1] ** 2 + b[2] ** 2)
        out.update(b1=b1, b2=b2, b3=b3)
    return out
