"""Shared fixtures for the parity tests (seeded synthetic inputs)."""
import itertools
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def seed42_array(nsub=None):
    """The synthetic array of /root/reference/tests/test_hessian_approx.py:73-102."""
    np.random.seed(42)
    npix, num_ants = 1024, 100
    pixsize = 0.5 * np.pi / 180 / 3600.0
    a1, a2 = np.asarray(list(itertools.combinations(range(num_ants), 2))).T
    antennas = 10e3 * np.random.normal(size=(num_ants, 3))
    antennas[:, 2] *= 0.001
    uvw = antennas[a1] - antennas[a2]
    freqs = np.linspace(700e6, 2000e6, 2)
    if nsub:
        uvw = uvw[::nsub]
    return npix, pixsize, uvw, freqs


def small_problem(nrow=400, nchan=3, nx=64, ny=48, seed=0, wscale=1.0, fov=0.25):
    """Random uvw filling the uv plane of an (nx, ny) image with a wide field (strong w-term)."""
    rng = np.random.default_rng(seed)
    freq = np.linspace(1.0e9, 1.2e9, nchan)
    cell = fov / max(nx, ny)
    umax = 0.5 / cell * 299792458.0 / freq.max()
    uvw = rng.uniform(-1, 1, (nrow, 3)) * umax * 0.95
    uvw[:, 2] *= 0.1 * wscale
    vis = rng.standard_normal((nrow, nchan)) + 1j * rng.standard_normal((nrow, nchan))
    wgt = rng.uniform(0.5, 1.5, (nrow, nchan))
    mask = (rng.uniform(size=(nrow, nchan)) > 0.05).astype(np.uint8)
    img = np.zeros((nx, ny))
    img[rng.integers(0, nx, 12), rng.integers(0, ny, 12)] = np.exp(rng.standard_normal(12))
    img[0, 0] = 1.0
    img[nx - 1, ny - 1] = 0.7  # image corners: worst case for the kernel error
    return dict(uvw=uvw, freq=freq, vis=vis, wgt=wgt, mask=mask, img=img, cell=cell, nx=nx, ny=ny)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) else np.float64)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(np.asarray(b).ravel()))
