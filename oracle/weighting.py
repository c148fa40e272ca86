"""TEST INFRASTRUCTURE ONLY — numpy restatement of the imaging-weight functions.

Follows /root/reference/src/pfb_imaging/utils/weighting.py:
  compute_counts      <- _compute_counts      (:81-140)
  counts_to_weights   <- counts_to_weights    (:143-208)
  filter_extreme_counts (:212-226), box_sum_counts (:229-254)
Pinned by tests/golden/weighting.npz, which was produced by executing those
reference functions (numba) in the build container (tests/golden/make_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this.
"""

import numpy as np
from scipy.ndimage import uniform_filter

LIGHTSPEED = 299792458.0


def uv_cells(uvw, freq, mask, nx, ny, cell_x, cell_y, usign=1.0, vsign=-1.0):
    """(nrow,nchan) int32 cell indices, -1 where masked or off the grid (weighting.py:103-135)."""
    u_cell = 1 / (nx * cell_x)
    umax = np.abs(1 / cell_x / 2)
    v_cell = 1 / (ny * cell_y)
    vmax = np.abs(1 / cell_y / 2)
    cn = np.asarray(freq, dtype=np.float64) / LIGHTSPEED
    u = np.asarray(uvw)[:, 0:1] * cn[None, :] * usign
    v = np.asarray(uvw)[:, 1:2] * cn[None, :] * vsign
    neg = v < 0
    u = np.where(neg, -u, u)
    v = np.where(neg, -v, v)
    ug = np.floor((u + umax) / u_cell)
    vg = np.floor((v + vmax) / v_cell)
    ok = (ug >= 0) & (ug < nx) & (vg >= 0) & (vg < ny)
    if mask is not None:
        ok &= np.asarray(mask) != 0
    ui = np.where(ok, ug, -1).astype(np.int32)
    vi = np.where(ok, vg, -1).astype(np.int32)
    return ui, vi


def compute_counts(uvw, freq, mask, wgt, nx, ny, cell_x, cell_y, dtype, ngrid=1, usign=1.0, vsign=-1.0):
    ui, vi = uv_cells(uvw, freq, mask, nx, ny, cell_x, cell_y, usign, vsign)
    ok = ui >= 0
    ncorr = wgt.shape[0]
    counts = np.zeros((ncorr, nx, ny), dtype=dtype)
    for c in range(ncorr):
        np.add.at(counts[c], (ui[ok], vi[ok]), wgt[c][ok])
    return counts


def counts_to_weights(counts, uvw, freq, weight, mask, nx, ny, cell_x, cell_y, robust, usign=1.0, vsign=-1.0):
    """In place on `weight` AND `counts`, like the reference."""
    if not counts.any():
        return weight
    ncorr = weight.shape[0]
    if robust > -2:
        numsqrt = 5 * 10 ** (-robust)
        avgwnum = (counts.astype(np.float64) ** 2).sum(axis=(1, 2))
        avgwden = counts.astype(np.float64).sum(axis=(1, 2))
        ssq = (numsqrt * numsqrt * avgwden / avgwnum).astype(weight.dtype)
        counts *= ssq[:, None, None]
        counts += 1
    ui, vi = uv_cells(uvw, freq, mask, nx, ny, cell_x, cell_y, usign, vsign)
    ok = ui >= 0
    for c in range(ncorr):
        cv = counts[c][ui[ok], vi[ok]]
        w = weight[c][ok]
        pos = cv > 0
        w[pos] = w[pos] / cv[pos]
        weight[c][ok] = w
    return weight


def filter_extreme_counts(counts, level=10.0):
    if not level:
        return counts
    ic, ix, iy = np.where(counts > 0)
    cnts = counts[ic, ix, iy]
    med = np.median(cnts)
    counts[ic, ix, iy] = np.maximum(cnts, med / level)
    return counts


def box_sum_counts(counts, npix_super):
    if npix_super is None or npix_super <= 0:
        return counts
    size = 2 * npix_super + 1
    out = np.empty_like(counts)
    for c in range(counts.shape[0]):
        out[c] = uniform_filter(counts[c], size=size, mode="constant", cval=0.0) * (size * size)
    return out


def l2_reweight(residual_vis, wgt, mask, dof, wgtp=None):
    """Student-t re-weighting of the natural weights from residual visibilities — restates the
    ``if l2_reweight_dof:`` block of image_data_products
    (/root/reference/src/pfb_imaging/operators/gridder.py:509-532); pinned by tests/golden/l2_reweight.npz
    (tests/golden/make_golden_l2.py executes that block itself).

    residual_vis (ncorr,nrow,nchan) complex, wgt (ncorr,nrow,nchan) real (a scaled COPY is returned),
    mask (nrow,nchan); returns None when the variance is zero (:531-532).  For ncorr > 1 the reference's
    ``if ovar:`` is ambiguous (numpy raises); this restatement continues when every ovar is non-zero."""
    p = 1.0 if wgtp is None else np.asarray(wgtp)
    ressq = (residual_vis * p * residual_vis.conj()).real  # :516
    sel = np.asarray(mask) > 0
    ssq = ressq[:, sel].sum(axis=-1)  # :519
    ovar = ssq / np.asarray(mask).sum()  # :520
    if not np.all(ovar):  # :521 (NaN counts as true, as in the reference)
        return None
    out = np.array(wgt, copy=True)
    out *= (dof + 2) / (dof + ressq / ovar[:, None, None])  # :527-530
    return out


def weight_data_corr(data, weight, jones, tbin_idx, tbin_counts, ant1, ant2):
    """numpy restatement of utils/correlations.py:195-232 (`_weight_data_impl`, `wgt_func`, `vis_func`): one
    correlation behind diagonal Jones terms, products in the reference's order."""
    data, weight, jones = np.asarray(data), np.asarray(weight), np.asarray(jones)
    nrow, nchan, _ = data.shape
    start = np.asarray(tbin_idx) - np.min(tbin_idx)
    row_t = np.full(nrow, -1)
    for t in range(start.size):
        row_t[start[t]:start[t] + tbin_counts[t]] = t
    ok = row_t >= 0
    vis = np.zeros((nrow, nchan), dtype=data.dtype)
    wgt = np.zeros((nrow, nchan), dtype=data.real.dtype)
    gp = jones[row_t[ok], np.asarray(ant1)[ok], :, 0, 0]
    gq = jones[row_t[ok], np.asarray(ant2)[ok], :, 0, 0]
    w0 = weight[ok][:, :, 0].astype(data.dtype)  # the real weight enters the products as a complex number
    v0 = data[ok][:, :, 0]
    wgt[ok] = (w0 * gp * gq * np.conjugate(gp) * np.conjugate(gq)).real
    vis[ok] = w0 * gq * v0 * np.conjugate(gp)
    return vis, wgt
