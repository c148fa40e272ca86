"""Seeded synthetic MeerKAT-like inputs for the benchmark configs (BASELINE.md §2,
SURVEY.md §8d).  The reference synthesises uvw with casacore
(``/root/reference/src/pfb_imaging/utils/astrometry.py:15-98``), which is absent
here; this is the standard ENU -> XYZ -> uvw rotation written out in numpy.
"""

from __future__ import annotations

import numpy as np

LIGHTSPEED = 299792458.0


def meerkat_like_antennas(rng, nant=64, ncore=48):
    """ENU positions (m): log-normal core (median 300 m, clipped 1 km) + outer ring 1-4 km."""
    r_core = np.minimum(np.exp(np.log(300.0) + 0.7 * rng.standard_normal(ncore)), 1000.0)
    r_out = rng.uniform(1000.0, 4000.0, nant - ncore)
    r = np.concatenate([r_core, r_out])
    th = rng.uniform(0.0, 2 * np.pi, nant)
    enu = np.stack([r * np.cos(th), r * np.sin(th), 5.0 * rng.standard_normal(nant)], axis=1)
    return enu


def uvw_tracks(enu, ntime, lat_deg=-30.71, dec_deg=-45.0, ha_range_h=(-3.0, 3.0)):
    """(nbl*ntime, 3) uvw in metres, time-major within baseline-major blocks."""
    lat, dec = np.deg2rad(lat_deg), np.deg2rad(dec_deg)
    e, n, u = enu[:, 0], enu[:, 1], enu[:, 2]
    # ENU -> equatorial XYZ
    x = -np.sin(lat) * n + np.cos(lat) * u
    y = e
    z = np.cos(lat) * n + np.sin(lat) * u
    xyz = np.stack([x, y, z], axis=1)
    i, j = np.triu_indices(enu.shape[0], k=1)
    bl = xyz[j] - xyz[i]  # (nbl, 3)
    ha = np.deg2rad(15.0 * np.linspace(ha_range_h[0], ha_range_h[1], ntime))
    sh, ch = np.sin(ha), np.cos(ha)
    sd, cd = np.sin(dec), np.cos(dec)
    uu = sh[:, None] * bl[None, :, 0] + ch[:, None] * bl[None, :, 1]
    vv = (-sd * ch)[:, None] * bl[None, :, 0] + (sd * sh)[:, None] * bl[None, :, 1] + cd * bl[None, :, 2]
    ww = (cd * ch)[:, None] * bl[None, :, 0] - (cd * sh)[:, None] * bl[None, :, 1] + sd * bl[None, :, 2]
    uvw = np.stack([uu, vv, ww], axis=2).reshape(-1, 3)  # (ntime*nbl, 3)
    return np.ascontiguousarray(uvw)


def band_freqs(band, nband, nchan, f_lo=856e6, f_hi=1712e6):
    edges = np.linspace(f_lo, f_hi, nband + 1)
    lo, hi = edges[band], edges[band + 1]
    return lo + (np.arange(nchan) + 0.5) * (hi - lo) / nchan


def default_cell(uvw, fmax, srf=2.0):
    """cell = 1/(2 b_max nu_max / c) / srf   (utils/misc.py:903-912)."""
    bmax = np.sqrt((uvw[:, :2] ** 2).sum(axis=1)).max()
    return 1.0 / (2.0 * bmax * fmax / LIGHTSPEED) / srf


def make_band(ntime, nchan, band=0, nband=8, seed=1234, precision="single", flag_frac=0.0, with_vis=True):
    """One imaging band of the benchmark: dict(uvw, freq, vis, wgt, mask)."""
    rng = np.random.default_rng(seed)
    enu = meerkat_like_antennas(rng)
    uvw = uvw_tracks(enu, ntime)
    freq = band_freqs(band, nband, nchan)
    nrow = uvw.shape[0]
    rdt = np.float32 if precision == "single" else np.float64
    cdt = np.complex64 if precision == "single" else np.complex128
    brng = np.random.default_rng([seed, band])
    out = dict(uvw=uvw, freq=freq)
    if with_vis:
        vis = np.empty((nrow, nchan), dtype=cdt)
        vis.real = brng.standard_normal((nrow, nchan), dtype=np.float32 if precision == "single" else np.float64)
        vis.imag = brng.standard_normal((nrow, nchan), dtype=np.float32 if precision == "single" else np.float64)
        out["vis"] = vis
    out["wgt"] = brng.uniform(0.5, 1.5, (nrow, nchan)).astype(rdt)
    if flag_frac > 0:
        out["mask"] = (brng.uniform(size=(nrow, nchan)) >= flag_frac).astype(np.uint8)
    else:
        out["mask"] = np.ones((nrow, nchan), dtype=np.uint8)
    return out


def point_source_image(nx, ny, nsrc=100, seed=7, dtype=np.float64):
    rng = np.random.default_rng(seed)
    img = np.zeros((nx, ny), dtype=dtype)
    ix = rng.integers(0, nx, nsrc)
    iy = rng.integers(0, ny, nsrc)
    img[ix, iy] = np.exp(rng.standard_normal(nsrc)).astype(dtype)
    return img
