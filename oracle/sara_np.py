"""TEST INFRASTRUCTURE ONLY — numpy restatement of the SARA backward-step pieces.

Follows the reference's in-tree numba code (all paths relative to /root/reference/src/pfb_imaging):
  * 1-D analysis / synthesis convolutions, pywt 'zero' mode: ``wavelets/convolutions.py:6-131, 134-330``
    (convention block at ``:315-338``):  y[i] = sum_k h[k] x[2i+1-k];  synthesis out[2t+p] = sum_m g[2m+p] c[t+K/2-1-m]
  * single / multi level 2-D transforms, x-first layout: ``wavelets/wavelets.py:209-343`` (`*_nocopyt`);
    the transposed layout of `dwt2d`/`idwt2d` (``:40-206``) is the same data transposed
  * dictionary: ``operators/psi.py:149-372`` (`PsiBand.dot/hdot`), ``:420-530`` (`PsiBandNocopyt`)
  * dual update / prox: ``prox/prox_21m.py:30-135``; primal-dual iteration ``opt/primal_dual.py:16-61, 404-448``

Pinned by tests/golden/sara.npz, which tests/golden/make_golden_sara.py produced by EXECUTING the reference's
own numba functions in this container (PyWavelets replaced by a stub serving the filters of
pfb_imaging_b200.wavelet_filters).  Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""

import numpy as np


def analysis_1d(x, h, axis):
    """y[i] = sum_k h[k] x[2i+1-k] along `axis`, zero outside, length (N+K-1)//2."""
    x = np.moveaxis(np.asarray(x, dtype=np.float64), axis, 0)
    N, K = x.shape[0], h.size
    nout = (N + K - 1) // 2
    out = np.zeros((nout,) + x.shape[1:])
    for i in range(nout):
        for k in range(K):
            n = 2 * i + 1 - k
            if 0 <= n < N:
                out[i] += h[k] * x[n]
    return np.moveaxis(out, 0, axis)


def synthesis_1d(c_lo, c_hi, g_lo, g_hi, nout, axis):
    """out[2t+p] = sum_m g_lo[2m+p] c_lo[t+K/2-1-m] + (same with hi), cropped to nout samples."""
    c_lo = np.moveaxis(np.asarray(c_lo, dtype=np.float64), axis, 0)
    c_hi = np.moveaxis(np.asarray(c_hi, dtype=np.float64), axis, 0)
    M, K = c_lo.shape[0], g_lo.size
    full = 2 * M - K + 2
    out = np.zeros((full,) + c_lo.shape[1:])
    for t in range(full // 2):
        for m in range(K // 2):
            i = t + K // 2 - 1 - m
            out[2 * t] += g_lo[2 * m] * c_lo[i] + g_hi[2 * m] * c_hi[i]
            out[2 * t + 1] += g_lo[2 * m + 1] * c_lo[i] + g_hi[2 * m + 1] * c_hi[i]
    return np.moveaxis(out[:nout], 0, axis)


def dwt2d(image, bk, b, fb):
    """Multi-level transform of `image` for basis b: (ntotx, ntoty) x-first coefficient block."""
    dec_lo, dec_hi = fb[0], fb[1]
    out = np.zeros((bk.ntotx[b], bk.ntoty[b]))
    a = np.asarray(image, dtype=np.float64)
    for l in range(bk.nlevel):
        sx, sy = int(bk.sx[b, l]), int(bk.sy[b, l])
        hx, hy = int(bk.ix[b, l, 1]), int(bk.iy[b, l, 1])
        lx, ly = hx - 2 * sx, hy - 2 * sy
        rows = np.concatenate([analysis_1d(a, dec_lo, 1), analysis_1d(a, dec_hi, 1)], axis=1)      # (nx_in, 2sy)
        blk = np.concatenate([analysis_1d(rows, dec_lo, 0), analysis_1d(rows, dec_hi, 0)], axis=0)  # (2sx, 2sy)
        out[lx:hx, ly:hy] = blk
        a = blk[:sx, :sy].copy()
    return out


def idwt2d(coeffs, bk, b, fb):
    """Inverse of dwt2d (exact for orthogonal filters), cropped to the image size."""
    rec_lo, rec_hi = fb[2], fb[3]
    alpha = np.array(coeffs, dtype=np.float64)
    img = None
    for l in range(bk.nlevel - 1, -1, -1):
        sx, sy = int(bk.sx[b, l]), int(bk.sy[b, l])
        hx, hy = int(bk.ix[b, l, 1]), int(bk.iy[b, l, 1])
        lx, ly = hx - 2 * sx, hy - 2 * sy
        if l < bk.nlevel - 1:
            alpha[lx:lx + sx, ly:ly + sy] = img[:sx, :sy]
        blk = alpha[lx:hx, ly:hy]
        nxo = min(int(bk.spx[b, l]), bk.nx)
        nyo = min(int(bk.spy[b, l]), bk.ny)
        cb = synthesis_1d(blk[:sx], blk[sx:], rec_lo, rec_hi, nxo, 0)          # (nxo, 2sy)
        img = synthesis_1d(cb[:, :sy], cb[:, sy:], rec_lo, rec_hi, nyo, 1)     # (nxo, nyo)
    out = np.zeros((bk.nx, bk.ny))
    out[:img.shape[0], :img.shape[1]] = img
    return out


def psi_dot(x, bk, fbs, transposed=False):
    """image (nx,ny) -> (nbasis, nxmax, nymax) [x-first] or (nbasis, nymax, nxmax) [transposed, `Psi`]."""
    out = np.zeros((bk.nbasis, bk.nxmax, bk.nymax))
    for b, name in enumerate(bk.bases):
        if name == "self":
            out[b, :bk.nx, :bk.ny] = x
        else:
            out[b, :bk.ntotx[b], :bk.ntoty[b]] = dwt2d(x, bk, b, fbs[b])
    return np.ascontiguousarray(out.transpose(0, 2, 1)) if transposed else out


def psi_hdot(alpha, bk, fbs, transposed=False):
    a = np.asarray(alpha, dtype=np.float64)
    if transposed:
        a = a.transpose(0, 2, 1)
    out = np.zeros((bk.nx, bk.ny))
    for b, name in enumerate(bk.bases):
        if name == "self":
            out += a[b, :bk.nx, :bk.ny]
        else:
            out += idwt2d(a[b, :bk.ntotx[b], :bk.ntoty[b]], bk, b, fbs[b])
    return out


def dual_update_fast(vp, v, lam, sigma, weight):
    """prox/prox_21m.py:104-135: v <- vtilde * min(1, lam w / |sum_band vtilde|), vtilde = vp + sigma v."""
    vt = vp + sigma * v
    s = np.abs(vt.sum(axis=0))
    thr = lam * weight
    scale = np.where(s > thr, thr / np.where(s > 0, s, 1.0), 1.0)
    return vt * scale[None]


def prox_21m(v, lam, sigma, weight):
    """prox/prox_21m.py:30-62 (prox_21m_numba): result = v * soft(|sum v / sigma|, lam w / sigma) / |sum| / sigma."""
    s = v.sum(axis=0) / sigma
    a = np.abs(s)
    soft = np.maximum(a - lam * weight / sigma, 0.0)
    ratio = np.where(s != 0, soft / np.where(a > 0, a, 1.0) / sigma, 0.0)
    return v * ratio[None]


def primal_dual(x, v, lam, psi_dot_f, psi_hdot_f, grad, hessnorm, nu, weight, tol, maxit, positivity=1, gamma=1.0):
    """opt/primal_dual.py:387-448 (`PrimalDual.solve` with the fused dual update)."""
    sigma = hessnorm / (2.0 * gamma) / nu
    tau = 0.98 / (hessnorm / (2.0 * gamma) + sigma * nu ** 2)
    x = x.copy(); v = v.copy()
    xp, vp = x.copy(), v.copy()
    eps, k = 1.0, 0
    for k in range(maxit):
        v = dual_update_fast(vp, psi_dot_f(xp), lam, sigma, weight)
        vp = 2.0 * v - vp
        xout = psi_hdot_f(vp) + grad(xp)
        x = xp - tau * xout
        if positivity == 1:
            x[x < 0] = 0.0
        elif positivity == 2:
            x[:, np.any(x <= 0, axis=0)] = 0.0
        eps = np.sqrt(((x - xp) ** 2).sum() / max((x ** 2).sum(), 1e-12)) if x.any() else 1.0
        if eps < tol:
            break
        xp, vp = x.copy(), v.copy()
    return x, v, k, eps


def forward_backward(x, lam, psi_dot_f, psi_hdot_f, grad, hessnorm, nu, weight, tol, maxit, positivity=0, gamma=1.0,
                     acceleration=True):
    """opt/forward_backward.py:84-133 (`ForwardBackward.solve` with the generic tight-frame prox)."""
    step = 2.0 * gamma / hessnorm
    x = x.copy()
    xp, y = x.copy(), x.copy()
    t, eps, k = 1.0, 1.0, 0
    for k in range(maxit):
        x = y - step * grad(y)
        a = psi_dot_f(x)
        x = x + psi_hdot_f(prox_21m(a, step * lam, 1.0, weight) - a) / nu
        if positivity == 1:
            x[x < 0] = 0.0
        elif positivity == 2:
            x[:, np.any(x <= 0, axis=0)] = 0.0
        eps = np.sqrt(((x - xp) ** 2).sum() / max((x ** 2).sum(), 1e-12)) if x.any() else 1.0
        if eps < tol:
            break
        if acceleration:
            tp = t
            t = (1.0 + np.sqrt(1.0 + 4.0 * tp ** 2)) / 2.0
            y = x + (tp - 1.0) / t * (x - xp)
        else:
            y = x.copy()
        xp = x.copy()
    return x, k, eps


def l1reweight(mcomps_sum, rmsfactor, rms_comps, alpha):
    """utils/misc.py:750-764 given the band-summed coefficients."""
    return (1 + rmsfactor) / (1 + np.abs(mcomps_sum) ** alpha / rms_comps[:, None, None] ** alpha)
