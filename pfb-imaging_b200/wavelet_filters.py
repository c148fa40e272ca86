"""Daubechies filter banks and the coefficient-packing bookkeeping of the SARA dictionary.

The reference obtains its filters from ``pywt.Wavelet(name).filter_bank`` (``operators/psi.py:56-61``);
PyWavelets is not part of this image, so db1..db5 are tabulated here.  The values are the
minimum-phase Daubechies scaling filters computed by spectral factorisation with 50-digit
arithmetic and rounded to double (PyWavelets' own table agrees with them to ~1e-13, its stored
precision).  Filter-bank conventions are PyWavelets':

    rec_lo = h,  dec_lo = h[::-1],  rec_hi[k] = (-1)^k h[K-1-k],  dec_hi = rec_hi[::-1]

The packing (`Bookkeeping`) restates ``operators/psi.py:24-137`` (`_build_wavelet_bookkeeping`).
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

_DB = {
    1: [0.7071067811865475244, 0.7071067811865475244],
    2: [0.48296291314453414337, 0.83651630373780790558, 0.22414386804201338103, -0.12940952255126038117],
    3: [0.332670552950082616, 0.80689150931109257649, 0.4598775021184915701, -0.1350110200102545887,
        -0.085441273882026661693, 0.035226291885709536603],
    4: [0.23037781330889650086, 0.71484657055291564709, 0.63088076792985890788, -0.027983769416859854211,
        -0.18703481171909308408, 0.030841381835560763627, 0.032883011666885199735, -0.010597401785069032105],
    5: [0.16010239797419291448, 0.60382926979718967054, 0.72430852843777292773, 0.13842814590132073151,
        -0.24229488706638203186, -0.032244869584638374648, 0.077571493840045713523, -0.0062414902127982742742,
        -0.012580751999081999469, 0.003335725285473771278],
}


def filter_bank(name: str):
    """(dec_lo, dec_hi, rec_lo, rec_hi) of 'db1'..'db5' as float64 arrays."""
    if not (name.startswith("db") and name[2:].isdigit() and int(name[2:]) in _DB):
        raise ValueError(f"unsupported wavelet {name!r}: this build tabulates db1..db5")
    h = np.array(_DB[int(name[2:])], dtype=np.float64)
    K = h.size
    rec_lo = h.copy()
    dec_lo = h[::-1].copy()
    rec_hi = np.array([(-1) ** k * h[K - 1 - k] for k in range(K)])
    dec_hi = rec_hi[::-1].copy()
    return dec_lo, dec_hi, rec_lo, rec_hi


def dwt_max_level(data_len: int, filter_len: int) -> int:
    """pywt.dwt_max_level: floor(log2(data_len / (filter_len - 1)))."""
    if filter_len < 2 or data_len < filter_len - 1:
        return 0
    return max(int(math.floor(math.log2(data_len / (filter_len - 1.0)))), 0)


def coeff_size(nsignal: int, nfilter: int) -> int:  # wavelets/wavelets.py:28-30
    return (nsignal + nfilter - 1) // 2


def signal_size(ncoeff: int, nfilter: int) -> int:  # wavelets/wavelets.py:33-35
    return 2 * ncoeff - nfilter + 2


@dataclass
class Bookkeeping:
    bases: tuple
    nlevel: int
    nx: int
    ny: int
    nbasis: int
    K: np.ndarray        # (nbasis,) filter length, 0 for 'self'
    ix: np.ndarray       # (nbasis, nlevel, 2) start/stop of the detail rows per level (x axis)
    iy: np.ndarray
    sx: np.ndarray       # (nbasis, nlevel) coefficient sizes
    sy: np.ndarray
    spx: np.ndarray      # (nbasis, nlevel) signal sizes reconstructed at each level
    spy: np.ndarray
    ntotx: np.ndarray    # (nbasis,)
    ntoty: np.ndarray
    nxmax: int
    nymax: int


def bookkeeping(nx: int, ny: int, bases, nlevel: int) -> Bookkeeping:
    bases = tuple(bases)
    nb = len(bases)
    K = np.zeros(nb, dtype=np.int64)
    ix = np.zeros((nb, nlevel, 2), dtype=np.int64)
    iy = np.zeros((nb, nlevel, 2), dtype=np.int64)
    sx = np.zeros((nb, nlevel), dtype=np.int64)
    sy = np.zeros((nb, nlevel), dtype=np.int64)
    spx = np.zeros((nb, nlevel), dtype=np.int64)
    spy = np.zeros((nb, nlevel), dtype=np.int64)
    ntotx = np.zeros(nb, dtype=np.int64)
    ntoty = np.zeros(nb, dtype=np.int64)
    nxmax, nymax = nx, ny
    for b, name in enumerate(bases):
        if name == "self":
            ntotx[b], ntoty[b] = nx, ny
            continue
        k = 2 * int(name[-1])  # operators/psi.py:75 (filter length from the name)
        filter_bank(name)
        K[b] = k
        if nlevel > dwt_max_level(min(nx, ny), k):
            raise ValueError(f"The requested decomposition level {nlevel} is not possible")
        cx_l, cy_l = [], []
        n_x, n_y = nx, ny
        tx = ty = 0
        for l in range(nlevel):
            cx, cy = coeff_size(n_x, k), coeff_size(n_y, k)
            cx_l.append(cx); cy_l.append(cy)
            sx[b, l], sy[b, l] = cx, cy
            spx[b, l], spy[b, l] = signal_size(cx, k), signal_size(cy, k)
            tx += cx; ty += cy
            n_x, n_y = cx + cx % 2, cy + cy % 2
        tx += cx_l[-1]; ty += cy_l[-1]
        ntotx[b], ntoty[b] = tx, ty
        nxmax, nymax = max(nxmax, tx), max(nymax, ty)
        lowx, lowy = cx_l[-1], cy_l[-1]
        ix[b, nlevel - 1] = (lowx, 2 * lowx)
        iy[b, nlevel - 1] = (lowy, 2 * lowy)
        lowx *= 2; lowy *= 2
        for l in reversed(range(nlevel - 1)):
            ix[b, l] = (lowx, lowx + cx_l[l])
            iy[b, l] = (lowy, lowy + cy_l[l])
            lowx += cx_l[l]; lowy += cy_l[l]
    return Bookkeeping(bases, nlevel, nx, ny, nb, K, ix, iy, sx, sy, spx, spy, ntotx, ntoty, int(nxmax), int(nymax))
