"""CPU: the host solvers against iterates produced by EXECUTING the reference's own pcg_numba / pcg /
power_method_numba / power_method (tests/golden/make_golden_solvers.py -> tests/golden/solvers.npz)."""
import os

import numpy as np

from pfb_imaging_b200 import solvers
from pfbg_testutil import GOLDEN

G = np.load(os.path.join(GOLDEN, "solvers.npz"))


def _aop():
    a = G["pcg_a"]
    return lambda v: (a @ v.ravel()).reshape(v.shape)


def test_pcg_iterates_match_reference_pcg_numba():
    aop, b, x0 = _aop(), G["pcg_b"], G["pcg_x0"]
    for k in (1, 2, 5, 12):
        x = solvers.pcg(aop, b, x0=x0.copy(), tol=0.0, maxit=k, minit=k, verbosity=0)
        np.testing.assert_allclose(x, G[f"pcg_numba_k{k}"], rtol=1e-10, atol=1e-12)
    x, r = solvers.pcg(aop, b, x0=None, tol=1e-9, maxit=400, minit=1, verbosity=0, return_resid=True)
    np.testing.assert_allclose(x, G["pcg_numba_conv"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(r, G["pcg_numba_conv_resid"], atol=1e-8)


def test_preconditioned_pcg_matches_reference_python_loop():
    aop, b, x0 = _aop(), G["pcg_b"], G["pcg_x0"]
    d = np.diag(G["pcg_a"]).reshape(b.shape)
    for k in (3, 9):
        x = solvers.pcg(aop, b, x0=x0.copy(), precond=lambda v: v / d, tol=0.0, maxit=k, minit=k, verbosity=0,
                        backtrack=False)
        np.testing.assert_allclose(x, G[f"pcg_k{k}"], rtol=1e-10, atol=1e-12)


def test_power_method_iterates_match_reference():
    aop = _aop()
    for k in (1, 4, 25):
        beta, bv = solvers.power_method(aop, G["pm_b0"].shape, b0=G["pm_b0"].copy(), tol=0.0, maxit=k, verbosity=0)
        np.testing.assert_allclose(beta, G[f"pm_numba_beta_k{k}"], rtol=1e-11)
        np.testing.assert_allclose(bv, G[f"pm_numba_b_k{k}"], rtol=1e-9, atol=1e-12)
    beta, bv = solvers.power_method(aop, G["pm_b0"].shape, b0=G["pm_b0"].copy(), tol=1e-10, maxit=3000, verbosity=0)
    np.testing.assert_allclose(beta, G["pm_beta_conv"], rtol=1e-9)
