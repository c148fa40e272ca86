"""CPU: the CG / power-method callers (solvers.py) on a synthetic SPD operator, single process and
band-sharded over 2 gloo ranks (scalar reductions through dist.allreduce_sum)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from pfb_imaging_b200 import solvers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spd(n, seed=0):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.linspace(1.0, 25.0, n)
    return (q * lam) @ q.T, lam


def test_pcg_solves_spd_system():
    a, _ = _spd(60)
    rng = np.random.default_rng(1)
    xt = rng.standard_normal((6, 10))
    b = (a @ xt.ravel()).reshape(6, 10)
    aop = lambda v: (a @ v.ravel()).reshape(v.shape)
    x0 = np.zeros_like(b)
    x, r = solvers.pcg(aop, b, x0=x0, tol=1e-12, maxit=300, minit=5, verbosity=0, return_resid=True)
    assert x is x0  # in place, like the reference
    np.testing.assert_allclose(x, xt, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r, aop(x) - b, atol=1e-8)
    # Jacobi preconditioner converges to the same solution
    d = np.diag(a).reshape(6, 10)
    x2 = solvers.pcg(aop, b, precond=lambda v: v / d, tol=1e-12, maxit=300, minit=5, verbosity=0)
    np.testing.assert_allclose(x2, xt, rtol=1e-8, atol=1e-9)
    # zero right-hand side: returns x0 untouched
    assert not solvers.pcg(aop, np.zeros_like(b), verbosity=0).any()


def test_power_method_finds_spectral_norm():
    a, lam = _spd(40, seed=3)
    beta, v = solvers.power_method(lambda z: (a @ z.ravel()).reshape(z.shape), (5, 8), tol=1e-10, maxit=2000,
                                   verbosity=0, seed=2)
    assert abs(beta - lam.max()) <= 1e-5 * lam.max()
    assert abs(np.linalg.norm(v) - 1.0) < 1e-12


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, os.environ["PFBG_ROOT"])
    from pfb_imaging_b200 import dist, solvers
    dist.init("gloo")
    r = dist.rank()
    nband, n = 4, 30
    rng = np.random.default_rng(0)
    mats = []
    for b in range(nband):
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        mats.append((q * np.linspace(1.0, 9.0 + b, n)) @ q.T)
    rhs = rng.standard_normal((nband, n))
    mine = dist.local_bands(nband)
    aop_loc = lambda v: np.stack([mats[b] @ v[i] for i, b in enumerate(mine)])
    x_loc = solvers.pcg(aop_loc, rhs[mine].copy(), tol=1e-13, maxit=200, minit=3, verbosity=0, reduce=dist.allreduce_sum)
    aop_all = lambda v: np.stack([mats[b] @ v[b] for b in range(nband)])
    x_all = solvers.pcg(aop_all, rhs.copy(), tol=1e-13, maxit=200, minit=3, verbosity=0)
    assert np.allclose(x_loc, x_all[mine], rtol=1e-9, atol=1e-11), np.abs(x_loc - x_all[mine]).max()
    b0 = rng.standard_normal((nband, n))
    beta_loc, _ = solvers.power_method(aop_loc, None, b0=b0[mine], tol=1e-11, maxit=3000, verbosity=0, reduce=dist.allreduce_sum)
    beta_all, _ = solvers.power_method(aop_all, None, b0=b0, tol=1e-11, maxit=3000, verbosity=0)
    assert abs(beta_loc - beta_all) < 1e-8 * beta_all
    print("RANK", r, "OK")
""")


def test_band_sharded_solvers_match_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PFBG_ROOT=ROOT, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RANK {r} OK" in o


def test_bench_host_logic():
    sys.path.insert(0, ROOT)
    import bench

    from pfb_imaging_b200.split import lpt_assign

    owner = lpt_assign([15.4, 17.0, 18.5, 19.2, 20.8, 22.4, 23.4, 25.6], 4)
    loads = [sum(c for b, c in enumerate([15.4, 17.0, 18.5, 19.2, 20.8, 22.4, 23.4, 25.6]) if owner[b] == r) for r in range(4)]
    assert sorted(set(owner)) == list(range(4)) and max(loads) - min(loads) < 2.0
    info = dict(nplanes=15, nu=6144, nv=6144, nx=4096, ny=4096)
    B = bench.algorithmic_bytes(info, 24998400, 16, 4)
    assert abs(B - (24998400 * (20 + 2 + 3) + 2 * 15 * (6 * 6144 * 6144 * 8 + 3 * 4096 * 4096 * 4))) < 1.0


def test_any_nonzero_matches_numpy_any():
    """The pool's zero-image short circuit (operators/hessian.py:47-48) scans bit patterns in chunks."""
    from pfb_imaging_b200.operators import _any_nonzero

    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64):
        z = np.zeros((37, 53), dt)
        assert not _any_nonzero(z) and not _any_nonzero(z, chunk=64)
        for pos in ((0, 0), (36, 52), (17, 3)):
            a = z.copy()
            a[pos] = rng.standard_normal() * 1e-30  # tiny but not zero
            assert _any_nonzero(a, chunk=100) == bool(a.any())
        assert _any_nonzero(np.full((3, 5), np.nan, dt))
    assert not _any_nonzero(np.zeros((0, 4)))


def test_band_pool_band_by_band_path_on_plain_operators():
    """Operators that are not device-pinned BandHessians go band by band; bands of other ranks stay zero."""
    from pfb_imaging_b200 import operators as ops

    class Scale:
        def __init__(self, f):
            self.f = f

        def dot(self, x):
            return self.f * x

    pool = ops.BandPool({0: Scale(2.0), 2: Scale(-1.0)}, nband=3)
    x = np.arange(3 * 4 * 5, dtype=np.float64).reshape(3, 4, 5)
    assert not pool._pipelined_ok(x)
    y = pool.hess_dot(x)
    assert np.array_equal(y[0], 2.0 * x[0]) and not y[1].any() and np.array_equal(y[2], -x[2])
