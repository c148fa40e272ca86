"""GPU: CG on the exact Hessian (the pcg_dds / fluxtractor pattern, /root/reference/src/pfb_imaging/opt/pcg.py:444-583)
and the BandPool facade (operators/band_worker.py:209-319)."""
from functools import partial

import numpy as np
import pytest

from pfb_imaging_b200 import operators as ops, solvers
from pfbg_testutil import rel_l2, small_problem

pytestmark = pytest.mark.gpu


def test_pcg_on_exact_hessian_recovers_model(gpu):
    p = small_problem(nrow=3000, nchan=2, nx=32, ny=32, seed=2)
    wgt = p["wgt"]
    wsum = float(wgt[p["mask"] != 0].sum())
    eta = 1e-3
    hess = partial(ops.hessian_slice, uvw=p["uvw"], weight=wgt, vis_mask=p["mask"], freq=p["freq"], beam=None,
                   cell=p["cell"], x0=0.0, y0=0.0, flip_u=False, flip_v=True, flip_w=False, do_wgridding=True,
                   epsilon=1e-9, double_accum=True, nthreads=1, eta=eta, wsum=wsum)
    model = np.zeros((32, 32))
    model[10, 12] = 2.0
    model[20, 5] = -1.0
    j = hess(model)  # rhs = H model, so CG must return the model
    x = solvers.pcg(hess, j, x0=np.zeros_like(j), tol=1e-9, maxit=150, minit=1, verbosity=0)
    assert rel_l2(hess(x), j) <= 1e-6
    assert rel_l2(x, model) <= 1e-3
    # power method: the spectral norm bounds the Rayleigh quotient of any vector
    beta, v = solvers.power_method(hess, (32, 32), tol=1e-6, maxit=200, verbosity=0, seed=0)
    rq = float(np.vdot(model, hess(model)) / np.vdot(model, model))
    assert beta >= rq * (1 - 1e-6) and beta > 0
    ops.clear_plan_cache()


def test_band_pool_roles(gpu):
    nband, nx = 3, 32
    probs = [small_problem(nrow=800, nchan=2, nx=nx, ny=nx, seed=10 + b) for b in range(nband)]
    bands = []
    for p in probs:
        wsum = float(p["wgt"][p["mask"] != 0].sum())
        bands.append(ops.BandHessian(p["uvw"], p["freq"], p["wgt"], p["mask"], nx, nx, p["cell"], epsilon=1e-8,
                                     eta=1e-2, wsum=wsum))
    pool = ops.BandPool(bands)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((nband, nx, nx))
    hx = pool.hess_dot(x)
    for b, p in enumerate(probs):
        wsum = float(p["wgt"][p["mask"] != 0].sum())
        ref = ops.hessian_slice(x[b], uvw=p["uvw"], weight=p["wgt"], vis_mask=p["mask"], freq=p["freq"], cell=p["cell"],
                                epsilon=1e-8, eta=1e-2, wsum=wsum)
        assert rel_l2(hx[b], ref) <= 1e-12  # tests/test_deconv.py:343-404: pool == in-process to 1e-12
    sol = pool.hess_cg(hx, tol=1e-10, maxit=200, minit=1)
    assert rel_l2(sol, x) <= 1e-4
    dirty = rng.standard_normal((nband, nx, nx))
    res0 = pool.residual(np.zeros_like(dirty), dirty)
    assert np.array_equal(res0, dirty)
    pool.close()
    ops.clear_plan_cache()
