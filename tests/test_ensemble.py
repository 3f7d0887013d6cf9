"""athena_b200.ensemble: several independent meshes stepped through HOST buffers concurrently (what
bench.py's e2e number runs).  A member that uploads its active zones, refreshes ghosts + W(U), does
one cycle and downloads again must hold exactly the state of a run that never left the device."""
import numpy as np
import pytest

import util
from athena_b200 import ensemble, pyhost


def _member_and_twin(lib, name, nsteps):
    meta, z = util.load_golden(name)
    B = (z["b1_0"], z["b2_0"], z["b3_0"])

    def make():
        ctx = util.make_ctx_from_global(lib, meta["mesh"], meta["block"], meta["xmin"], meta["xmax"], meta["ng"],
                                        True, meta["gamma"], z["u0"], B, flux=meta.get("flux"))
        ti = pyhost.TimeIntegrator(ctx, meta["integrator"], meta["xorder"], meta["cfl"], fused=True)
        ti.initialize()
        return ctx, ti
    ctx, ti = make()
    twin, tit = make()
    for _ in range(nsteps):
        tit.step()
    nb = len(ctx.gids)
    n1, n2, n3 = ctx.mesh.block_nx
    shp = {"U": (5, n3, n2, n1), "B_X1F": (n3, n2, n1 + 1), "B_X2F": (n3, n2 + 1, n1), "B_X3F": (n3 + 1, n2, n1)}
    host = {k: np.empty((nb,) + s) for k, s in shp.items()}
    m = ensemble.Member(ctx, ti, host)
    m.download()   # the host copy starts as the device state
    return m, twin


def _check(lib):
    nsteps = 3
    pairs = [_member_and_twin(lib, "mhd_linwave_vl2_plm_hlld", nsteps) for _ in range(3)]
    ens = ensemble.HostSteppedEnsemble([p[0] for p in pairs])
    ens.run(nsteps, lag_s=0.001)
    ens.sync()
    for m, twin in pairs:
        U, B = util.gather_from_ctx(m.ctx)
        Ut, Bt = util.gather_from_ctx(twin)
        np.testing.assert_array_equal(U, Ut)
        for a, b in zip(B, Bt):
            np.testing.assert_array_equal(a, b)
        assert m.ti.ncycle == nsteps
        # and the member's host buffers hold that state (active zones, block by block)
        ng = m.ctx.mesh.ng
        u0 = m.ctx.download("U", 0)
        np.testing.assert_array_equal(m.host["U"][0], u0[:, ng:-ng, ng:-ng, ng:-ng])
        m.ctx.close()
        twin.close()


def test_ensemble_members_through_host_buffers_match_device_resident_runs(emul_lib):
    _check(emul_lib)


@pytest.mark.gpu
def test_ensemble_members_through_host_buffers_match_device_resident_runs_on_gpu(cuda_lib):
    _check(cuda_lib)
