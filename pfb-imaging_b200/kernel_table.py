"""ES-kernel parameter table for the w-gridder plan.

The gridding kernel is the exponential-of-semicircle ("ES") window
``phi(x) = exp(beta*(sqrt(1-x^2)-1))`` on ``|x|<1`` — the same definition the
reference carries in-tree (``src/pfb_imaging/utils/weighting.py:25-35``) and
ducc0.wgridder 0.41.0 (not in tree) uses in a generalised form.

For a support of ``W`` cells and an oversampling factor ``sigma`` the table
stores the shape parameter ``beta`` that minimises the worst-case 1-D
interpolation error

    err(W, beta, sigma) = max_{|xi| <= 1/(2 sigma)}  rms_{g in [0,1)}
        | exp(2 pi i g xi) - psihat(xi)^-1 sum_j psi(g-j) exp(2 pi i j xi) |

with ``psi(t) = phi(2t/W)`` and ``psihat`` its Fourier transform.  The plan
(``plan.py``) picks the smallest ``W`` whose error is below the requested
epsilon (split over the 2 or 3 interpolated axes).

Run ``python kernel_table.py`` to regenerate ``kernel_table.json``.
"""

from __future__ import annotations

import json
import os

import numpy as np
from numpy.polynomial.legendre import leggauss

W_MIN, W_MAX = 4, 16
SIGMAS = [round(1.10 + 0.05 * i, 2) for i in range(31)]  # 1.10 .. 2.60

_GL_X, _GL_W = leggauss(160)


def es_kernel(x, beta):
    """phi(x) = exp(beta*(sqrt(1-x^2)-1)) for |x|<=1, else 0."""
    x = np.asarray(x, dtype=np.float64)
    arg = (1.0 - x) * (1.0 + x)
    inside = arg >= 0.0
    out = np.exp(beta * (np.sqrt(np.where(inside, arg, 0.0)) - 1.0))
    return np.where(inside, out, 0.0)


def kernel_ft(xi, W, beta):
    """psihat(xi) = (W/2) * int_{-1}^{1} phi(x) cos(pi W xi x) dx  (Gauss-Legendre, fp64)."""
    xi = np.atleast_1d(np.asarray(xi, dtype=np.float64))
    ph = es_kernel(_GL_X, beta) * _GL_W
    return 0.5 * W * (np.cos(np.pi * W * np.outer(xi, _GL_X)) @ ph)


def kernel_error(W, beta, sigma, nxi=48, ng=24):
    """Worst-case (over image position) rms (over sub-cell offset) 1-D error."""
    xi = np.linspace(0.0, 0.5 / sigma, nxi)
    g = (np.arange(ng) + 0.5) / ng
    j0 = np.floor(g - 0.5 * W).astype(np.int64) + 1
    j = j0[:, None] + np.arange(W)[None, :]
    psi = es_kernel(2.0 * (g[:, None] - j) / W, beta)
    ph = np.exp(2j * np.pi * j[None, :, :] * xi[:, None, None])
    approx = (psi[None] * ph).sum(axis=2) / kernel_ft(xi, W, beta)[:, None]
    exact = np.exp(2j * np.pi * g[None, :] * xi[:, None])
    e = np.abs(approx - exact)
    return float(np.sqrt((e * e).mean(axis=1)).max())


def optimise_beta(W, sigma):
    b0 = np.pi * W * (1.0 - 0.5 / sigma)
    best = (np.inf, None)
    for gam in np.linspace(0.85, 1.05, 81):
        e = kernel_error(W, gam * b0, sigma)
        if e < best[0]:
            best = (e, gam * b0)
    # local refinement
    lo, hi = best[1] - 0.0025 * b0, best[1] + 0.0025 * b0
    for b in np.linspace(lo, hi, 9):
        e = kernel_error(W, b, sigma)
        if e < best[0]:
            best = (e, b)
    return best


def build_table():
    tab = {}
    for s in SIGMAS:
        for W in range(W_MIN, W_MAX + 1):
            err, beta = optimise_beta(W, s)
            tab[f"{s:.2f}:{W}"] = [float(beta), float(err)]
    return tab


_TABLE = None


def load_table():
    global _TABLE
    if _TABLE is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_table.json")
        with open(path) as f:
            _TABLE = json.load(f)
    return _TABLE


def lookup(sigma, W):
    """(beta, err) valid for any oversampling >= the tabulated sigma <= `sigma`."""
    tab = load_table()
    s_tab = max([s for s in SIGMAS if s <= sigma + 1e-12], default=None)
    if s_tab is None:
        raise ValueError(f"oversampling {sigma} below the tabulated minimum {SIGMAS[0]}")
    beta, err = tab[f"{s_tab:.2f}:{W}"]
    return beta, err


if __name__ == "__main__":
    t = build_table()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_table.json")
    with open(out, "w") as f:
        json.dump(t, f, indent=0, sort_keys=True)
    print("wrote", out, len(t), "entries")
