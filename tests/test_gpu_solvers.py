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


def test_band_pool_pipelined_matches_band_by_band(gpu, monkeypatch):
    """hess_dot overlaps the copies of neighbouring bands with the kernels: same numbers as band by band,
    for pageable and page-locked (second sighting) input cubes, a zero band and a band owned by another rank."""
    nband, nx, ny = 6, 384, 352  # 1.08 MB per band in fp64: above the page-locking threshold
    probs = {b: small_problem(nrow=600, nchan=2, nx=nx, ny=ny, seed=20 + b) for b in (0, 1, 3, 4, 5)}  # band 2: not ours
    bands = {}
    for b, p in probs.items():
        wsum = float(p["wgt"][p["mask"] != 0].sum())
        bands[b] = ops.BandHessian(p["uvw"], p["freq"], p["wgt"], p["mask"], nx, ny, p["cell"], epsilon=1e-8, eta=1e-2,
                                   wsum=wsum)
    pool = ops.BandPool(bands, nband=nband)
    x = np.random.default_rng(5).standard_normal((nband, nx, ny))
    x[1] = 0.0
    monkeypatch.setenv("PFBG_POOL_PIPELINE", "0")
    ref = pool.hess_dot(x)
    monkeypatch.setenv("PFBG_POOL_PIPELINE", "1")
    outs = [pool.hess_dot(x) for _ in range(3)]  # pageable, then page-locked input
    for o in outs:
        assert o.shape == x.shape and o.dtype == x.dtype
        assert not o[1].any() and not o[2].any()
        for b in (0, 3, 4, 5):  # bands 4 and 5 run concurrently on the two compute streams
            # sigma = 1.25 / W = 16 at this size: two applies of the SAME path already differ by ~5e-10 (the order of
            # the grid atomics, amplified by 1 / psihat at the image edge); the bound is epsilon
            assert rel_l2(o[b], ref[b]) <= 1e-8
    assert outs[0] is not outs[1] and not np.shares_memory(outs[0], outs[1])  # every call returns its own array
    y = np.random.default_rng(6).standard_normal((nband, nx, ny)).astype(np.float32)  # wrong dtype: band-by-band path
    assert rel_l2(pool.hess_dot(y)[0], pool.hess_dot(y.astype(np.float64))[0]) <= 1e-6
    pool.close()
    ops.clear_plan_cache()


def test_device_resident_cg_matches_host_cg(gpu):
    """BandHessian.cg keeps every CG vector on the device; same iterates as solvers.pcg on the numpy operator."""
    p = small_problem(nrow=2500, nchan=2, nx=48, ny=40, seed=4)
    wsum = float(p["wgt"][p["mask"] != 0].sum())
    rng = np.random.default_rng(1)
    beam = rng.uniform(0.5, 1.0, (48, 40))
    op = ops.BandHessian(p["uvw"], p["freq"], p["wgt"], p["mask"], 48, 40, p["cell"], beam=beam, epsilon=1e-9, eta=5e-3,
                         wsum=wsum)
    model = np.zeros((48, 40))
    model[10, 12], model[30, 7] = 2.0, -1.0
    rhs = op.dot(model)
    for maxit in (7, 150):  # a fixed number of iterations (same iterates) and a converged solve
        a = solvers.pcg(op.dot, rhs, x0=np.zeros_like(rhs), tol=1e-10, maxit=maxit, minit=1, verbosity=0)
        b = op.cg(rhs, tol=1e-10, maxit=maxit, minit=1)
        assert rel_l2(b, a) <= 1e-7
    assert rel_l2(b, model) <= 1e-3
    x0 = 0.5 * model
    c = op.cg(rhs, x0=x0, tol=1e-10, maxit=150, minit=1)
    assert rel_l2(c, model) <= 1e-3
    assert not op.cg(np.zeros_like(rhs), tol=1e-10, maxit=5).any()  # zero right-hand side: "initial residual is zero"
    op.close()


def test_device_power_method_matches_host(gpu):
    p = small_problem(nrow=1500, nchan=2, nx=32, ny=40, seed=6)
    wsum = float(p["wgt"][p["mask"] != 0].sum())
    op = ops.BandHessian(p["uvw"], p["freq"], p["wgt"], p["mask"], 32, 40, p["cell"], epsilon=1e-8, eta=1e-2, wsum=wsum)
    b0 = np.random.default_rng(3).standard_normal((32, 40))
    beta_h, v_h = solvers.power_method(op.dot, (32, 40), b0=b0, tol=1e-8, maxit=300, verbosity=0)
    beta_d, v_d = solvers.power_method_device(op.dot_dev, (32, 40), b0=b0, tol=1e-8, maxit=300, verbosity=0)
    assert abs(beta_d - beta_h) <= 1e-6 * abs(beta_h)
    assert min(rel_l2(v_d, v_h), rel_l2(v_d, -v_h)) <= 1e-3
    beta_s, _ = op.spectral_norm(tol=1e-8, maxit=300, seed=0)
    assert abs(beta_s - beta_h) <= 1e-2 * abs(beta_h)  # other start vector; the stopping rule is on the change of beta
    op.close()
