"""Callers of the Hessian on the hot path: conjugate gradients and the power method, with the
scalar reductions routed through an optional cross-band all-reduce.

Counterparts in the reference (/root/reference/src/pfb_imaging):
  pcg            opt/pcg.py:202-314  (python-loop CG; `pcg_numba` :88-199 is the fused variant)
  power_method   opt/power_method.py:40-92
  norm_diff      opt/pcg.py:70-85
The iteration contracts are kept: `eps = ||x - x_prev|| / ||x||`, at least `minit` and at most
`maxit` iterations, stop after 5 stalled iterations; the power method stops on the relative change of
the Rayleigh quotient.  When the cube is band-sharded (each rank holds its own bands, dist.py), pass
`reduce=dist.allreduce_sum`: every dot product / norm then becomes ONE small all-reduce
(2-3 doubles per iteration), which is all the communication these loops need.
"""

from __future__ import annotations

import numpy as np


def _identity(v):
    return v


def _dots(reduce, *pairs):
    """Real parts of <a,b> for each pair, summed over ranks with one message."""
    loc = np.array([float(np.vdot(a, b).real) for a, b in pairs], dtype=np.float64)
    return reduce(loc) if reduce is not None else loc


def norm_diff(x, xp, reduce=None):
    d2, x2 = _dots(reduce, (x - xp, x - xp), (x, x))
    return float(np.sqrt(d2 / x2)) if x2 > 0 else 0.0


def pcg(aop, b, x0=None, precond=None, tol=1e-5, maxit=500, minit=100, verbosity=1, report_freq=10, backtrack=True,
        return_resid=False, reduce=None):
    """Preconditioned conjugate gradients for ``aop(x) = b`` (aop symmetric positive definite).

    `x0` is updated in place and returned, like the reference."""
    x = np.zeros(b.shape, dtype=b.dtype) if x0 is None else x0
    minv = _identity if precond is None else precond
    r = aop(x) - b
    y = minv(r)
    if not _dots(reduce, (y, y))[0] > 0.0:
        if verbosity:
            print("Initial residual is zero")
        return (x, r) if return_resid else x
    p = -y
    rho = _dots(reduce, (r, y))[0]
    phi0 = rho if (np.isfinite(rho) and rho != 0.0) else 1.0
    k, eps, stalls = 0, 1.0, 0
    xprev = np.empty_like(x)
    while (eps > tol or k < minit) and k < maxit and stalls < 5:
        np.copyto(xprev, x)
        ap = aop(p)
        rho, pap = _dots(reduce, (r, y), (p, ap))
        if rho == 0.0 and pap == 0.0:
            # the residual cancelled EXACTLY (small masked systems converge in a couple of iterations and the loop
            # runs on until ||x - xprev|| drops): the reference forms 0 / 0 here and returns NaNs; the iterate is
            # already the solution
            break
        alpha = rho / pap
        x += alpha * p
        r = r + alpha * ap
        y = minv(r)
        rho_next = _dots(reduce, (r, y))[0]
        p *= rho_next / rho
        p -= y
        k += 1
        eps_prev, eps = eps, norm_diff(x, xprev, reduce)
        if abs(eps_prev - eps) < 1e-3 * tol:
            stalls += 1
        if verbosity > 1 and k % report_freq == 0:
            print(f"At iteration {k} eps = {eps:.3e}, phi = {rho_next / phi0:.3e}")
    if verbosity:
        if k >= maxit:
            print(f"Max iters reached. eps = {eps:.3e}")
        elif stalls >= 5:
            print(f"Stalled after {k} iterations with eps = {eps:.3e}")
        else:
            print(f"Success, converged after {k} iterations")
    return (x, r) if return_resid else x


def power_method(aop, imsize, b0=None, tol=1e-5, maxit=250, verbosity=1, report_freq=25, reduce=None, seed=None):
    """Largest eigenvalue (spectral norm) of a symmetric operator and its eigenvector."""
    if b0 is None:
        b = np.random.default_rng(seed).standard_normal(imsize)
    else:
        b = np.array(b0, dtype=np.float64, copy=True)
    b /= np.sqrt(_dots(reduce, (b, b))[0])
    beta, eps, k = 1.0, 1.0, 0
    while eps > tol and k < maxit:
        ab = aop(b)
        num, den, nrm2 = _dots(reduce, (b, ab), (b, b), (ab, ab))
        beta_prev, beta = beta, num / den
        b = ab / np.sqrt(nrm2)
        eps = abs(beta - beta_prev) / abs(beta_prev)
        k += 1
        if verbosity > 1 and k % report_freq == 0:
            print(f"At iteration {k} eps = {eps:.3e}")
    if verbosity:
        print(f"Maximum iterations reached. eps = {eps:.3e}, beta = {beta:.3e}" if k == maxit
              else f"Success, converged after {k} iterations. beta = {beta:.3e}")
    return beta, b


def pcg_device(apply_dev, b, x0=None, tol=1e-5, maxit=500, minit=100, verbosity=1, report_freq=10, device=0, reduce=None):
    """Conjugate gradients with every vector resident on the device (identity preconditioner).

    Same iteration and stopping contract as `pcg` (``opt/pcg.py:202-314``).  `apply_dev(in_ptr, out_ptr, stream)`
    applies the operator to device arrays (e.g. ``GridderPlan.hessian_dev`` with the band's beam / wsum / eta bound);
    `b` / `x0` are numpy arrays, the solution comes back as numpy.  Per iteration: one operator apply, three axpby
    kernels and two fused dot-product kernels; ``eps = ||x - x_prev|| / ||x||`` is formed as ``|alpha| ||p|| / ||x||``
    so no copy of the previous iterate is kept.  `reduce` sums the scalar pairs over ranks for band-sharded cubes.
    Like the reference's solvers (``opt/pcg.py:106-112``) a given `x0` of the operator's dtype is bound as the
    iterate: it receives the solution and IS the returned array."""
    import ctypes as C

    import torch

    from . import _lib

    lib = _lib.load()
    rdt = np.dtype(b.dtype)
    prec = _lib.PFBG_F32 if rdt == np.float32 else _lib.PFBG_F64
    dev = torch.device("cuda", device)
    bt = torch.from_numpy(np.ascontiguousarray(b)).to(dev)
    x = torch.zeros_like(bt) if x0 is None else torch.from_numpy(np.ascontiguousarray(x0, dtype=rdt)).to(dev)
    r, p, ap = torch.empty_like(bt), torch.empty_like(bt), torch.empty_like(bt)
    s = torch.cuda.current_stream(dev).cuda_stream
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    n = bt.numel()
    out2 = (C.c_double * 2)()

    def result():
        res = x.cpu().numpy()
        if isinstance(x0, np.ndarray) and x0.dtype == rdt and x0.shape == res.shape and x0.flags.writeable:
            x0[...] = res
            return x0
        return res

    def axpby(out, a, xx, bb, yy):
        _lib.check(lib.pfbs_axpby(prec, device, ptr(out), float(a), ptr(xx), float(bb), ptr(yy), n, s))

    def dot2(a, bq, c, d):
        _lib.check(lib.pfbs_dot2(prec, device, ptr(a), ptr(bq), ptr(c), ptr(d), n, out2, s))
        loc = np.array([out2[0], out2[1]])
        return reduce(loc) if reduce is not None else loc

    apply_dev(ptr(x), ptr(r), s)
    axpby(r, 1.0, r, -1.0, bt)          # r = A x - b
    rho = dot2(r, r, r, r)[0]
    if not rho > 0.0:
        if verbosity:
            print("Initial residual is zero")
        return result()
    axpby(p, -1.0, r, 0.0, r)           # p = -r
    phi0 = rho if np.isfinite(rho) else 1.0
    k, eps, stalls = 0, 1.0, 0
    while (eps > tol or k < minit) and k < maxit and stalls < 5:
        apply_dev(ptr(p), ptr(ap), s)
        pap, pp = dot2(p, ap, p, p)
        if rho == 0.0 and pap == 0.0:  # exact convergence: see pcg
            break
        alpha = rho / pap
        axpby(x, 1.0, x, alpha, p)
        axpby(r, 1.0, r, alpha, ap)
        rho_next, xx = dot2(r, r, x, x)
        axpby(p, rho_next / rho, p, -1.0, r)
        rho = rho_next
        k += 1
        eps_prev, eps = eps, (abs(alpha) * np.sqrt(pp / xx) if xx > 0 else 0.0)
        if abs(eps_prev - eps) < 1e-3 * tol:
            stalls += 1
        if verbosity > 1 and k % report_freq == 0:
            print(f"At iteration {k} eps = {eps:.3e}, phi = {rho / phi0:.3e}")
    if verbosity:
        if k >= maxit:
            print(f"Max iters reached. eps = {eps:.3e}")
        elif stalls >= 5:
            print(f"Stalled after {k} iterations with eps = {eps:.3e}")
        else:
            print(f"Success, converged after {k} iterations")
    return result()


def power_method_device(apply_dev, shape, dtype=np.float64, b0=None, tol=1e-5, maxit=250, verbosity=1, report_freq=25,
                        device=0, reduce=None, seed=None):
    """Power iteration with the vector resident on the device (``opt/power_method.py:40-92`` contract: stop on the
    relative change of the Rayleigh quotient).  `apply_dev(in_ptr, out_ptr, stream)` as in `pcg_device`.
    Returns ``(beta, b)`` with `b` a numpy array."""
    import ctypes as C

    import torch

    from . import _lib

    lib = _lib.load()
    rdt = np.dtype(dtype)
    prec = _lib.PFBG_F32 if rdt == np.float32 else _lib.PFBG_F64
    dev = torch.device("cuda", device)
    b = np.random.default_rng(seed).standard_normal(shape) if b0 is None else np.asarray(b0)
    bt = torch.from_numpy(np.ascontiguousarray(b, dtype=rdt)).to(dev)
    ab = torch.empty_like(bt)
    s = torch.cuda.current_stream(dev).cuda_stream
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    n = bt.numel()
    out2 = (C.c_double * 2)()

    def dot2(a, bq, c, d):
        _lib.check(lib.pfbs_dot2(prec, device, ptr(a), ptr(bq), ptr(c), ptr(d), n, out2, s))
        loc = np.array([out2[0], out2[1]])
        return reduce(loc) if reduce is not None else loc

    nrm2 = dot2(bt, bt, bt, bt)[0]
    _lib.check(lib.pfbs_axpby(prec, device, ptr(bt), 1.0 / np.sqrt(nrm2), ptr(bt), 0.0, ptr(bt), n, s))
    beta, eps, k = 1.0, 1.0, 0
    while eps > tol and k < maxit:
        apply_dev(ptr(bt), ptr(ab), s)
        num, nrm2 = dot2(bt, ab, ab, ab)  # b is unit norm: the Rayleigh quotient is <b, A b>
        beta_prev, beta = beta, num
        _lib.check(lib.pfbs_axpby(prec, device, ptr(bt), 1.0 / np.sqrt(nrm2), ptr(ab), 0.0, ptr(ab), n, s))
        eps = abs(beta - beta_prev) / abs(beta_prev)
        k += 1
        if verbosity > 1 and k % report_freq == 0:
            print(f"At iteration {k} eps = {eps:.3e}")
    if verbosity:
        print(f"Maximum iterations reached. eps = {eps:.3e}, beta = {beta:.3e}" if k == maxit
              else f"Success, converged after {k} iterations. beta = {beta:.3e}")
    return beta, bt.cpu().numpy()


def pcg_dds(ds_name, eta, mask=1.0, use_psf=True, residual_name="RESIDUAL", model_name="MODEL", do_wgridding=True,
            epsilon=5e-4, double_accum=True, nthreads=1, zero_model_outside_mask=False, tol=1e-5, maxit=500,
            verbosity=1, report_freq=10):
    """Flux mop of one band: CG on the EXACT Hessian ``beam R^H W R beam / wsum + eta`` (``opt/pcg.py:444-583``, the
    one place the reference iterates on the gridder Hessian; `pfb fluxtractor`).  `ds_name` is the band's ``.dds``
    zarr group (path, list holding one, or an opened dataset-like).  Every CG iteration is one fused device apply
    of the band pinned by the operator-level plan cache; MODEL_MOPPED / RESIDUAL_MOPPED / UPDATE / X0 are written
    back to the store.  Returns ``(resid, bandid)``."""
    from functools import partial

    from .operators import _attr, _open_dataset, hessian_slice

    if isinstance(ds_name, (list, tuple)):
        ds_name = ds_name[0]
    ds = _open_dataset(ds_name, drop_vars=["PSF", "PSFHAT"])
    geom = dict(uvw=ds.UVW.values, weight=ds.WEIGHT.values, vis_mask=ds.MASK.values, freq=ds.FREQ.values,
                cell=_attr(ds, "cell_rad"), x0=_attr(ds, "x0"), y0=_attr(ds, "y0"), do_wgridding=do_wgridding,
                epsilon=epsilon, double_accum=double_accum, nthreads=nthreads)
    beam0 = ds.BEAM.values
    beam = mask * beam0
    if zero_model_outside_mask:
        if model_name not in ds:
            raise RuntimeError(f"Asked to zero model outside mask but {model_name} not in dds")
        model = np.where(mask > 0, ds[model_name].values, 0.0)
        print("Zeroing model outside mask")
        resid = ds.DIRTY.values - hessian_slice(model, beam=beam0, **geom)
        j = resid * beam
    else:
        model = np.array(ds[model_name].values, copy=True) if model_name in ds else np.zeros(np.shape(mask), dtype=float)
        if residual_name in ds:
            j = ds[residual_name].values * beam
            ds = ds.drop_vars(residual_name)
        else:
            j = ds.DIRTY.values * beam
    wsum = _attr(ds, "wsum")
    j = j / wsum
    x0 = ds.UPDATE.values * mask if "UPDATE" in ds else np.zeros_like(j)
    hess = partial(hessian_slice, beam=np.ascontiguousarray(beam), flip_u=_attr(ds, "flip_u"), flip_v=_attr(ds, "flip_v"),
                   flip_w=_attr(ds, "flip_w"), eta=eta, wsum=wsum, **geom)
    x = pcg(hess, j, x0=np.array(x0, copy=True), precond=None, tol=tol, maxit=maxit, minit=1, verbosity=verbosity,
            report_freq=report_freq, backtrack=False, return_resid=False)
    model = model + x
    resid = ds.DIRTY.values - hessian_slice(model, beam=beam0, **geom)
    if hasattr(ds, "assign") and hasattr(ds, "to_zarr"):
        out = ds.assign(MODEL_MOPPED=(("x", "y"), model), RESIDUAL_MOPPED=(("x", "y"), resid), UPDATE=(("x", "y"), x),
                        X0=(("x", "y"), x0))
        target = ds_name if isinstance(ds_name, (str, bytes)) else getattr(ds, "_path", None)
        if target is not None:
            out[["MODEL_MOPPED", "RESIDUAL_MOPPED", "UPDATE", "X0"]].to_zarr(target, mode="a")
    return resid, int(_attr(ds, "bandid", 0))
