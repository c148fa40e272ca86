"""Device SARA backward step: the wavelet dictionary, the l21 regulariser and the
primal-dual loop, as drop-ins for (/root/reference/src/pfb_imaging):

  Psi / PsiNocopyt      operators/psi.py:549-664   (.dot(x, alphao) / .hdot(alpha, xo), in-place, numpy)
  L21                   prox/l21.py:17-91          (.prox / .dual_update / .l1weight)
  PrimalDual            opt/primal_dual.py:284-448 (.setup / .set_grad / .solve)

numpy in / numpy out for the reference-facing methods; `PrimalDual.solve` keeps the primal and dual cubes on
the device for the whole loop (torch tensors are used as device buffers only).  The transforms and updates
run in ``libpfbgrid.so`` (``csrc/sara.cuh``, C ABI ``include/pfbsara.h``).  When the bands of a cube are
sharded over ranks (dist.py) the l21 band sum and the convergence norm become one all-reduce each per
iteration: pass ``reduce=dist.allreduce_sum_tensor``-style callables (see `PrimalDual`).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import wavelet_filters as wf
from .wgridder import current_device

_PREC = {np.dtype(np.float32): _lib.PFBG_F32, np.dtype(np.float64): _lib.PFBG_F64}
TRANSPOSED = 8  # PFBS_TRANSPOSED


def _vp(a):
    return C.c_void_p(a.ctypes.data)


class PsiNocopyt:
    """SARA dictionary with x-first coefficient cubes ``(nband, nbasis, nxmax, nymax)``."""

    _transposed = False

    def __init__(self, nband, nx, ny, bases, nlevel, nthreads=1, dtype=np.float64, device=None):
        self.nband, self.nx, self.ny = int(nband), int(nx), int(ny)
        self.bases = tuple(bases)
        self.nbasis = len(self.bases)
        self.nlevel = int(nlevel)
        self.nthreads = nthreads  # accepted for signature compatibility; the device does the work
        self.rdt = np.dtype(dtype)
        if self.rdt not in _PREC:
            raise TypeError("dtype must be float32 or float64")
        self.prec = _PREC[self.rdt]
        self.device = current_device() if device is None else int(device)
        self.bk = bk = wf.bookkeeping(self.nx, self.ny, self.bases, self.nlevel)
        self.nxmax, self.nymax = bk.nxmax, bk.nymax
        filt = np.zeros((self.nbasis, 4, 10), dtype=np.float64)
        for b, name in enumerate(self.bases):
            if name != "self":
                for t, f in enumerate(wf.filter_bank(name)):
                    filt[b, t, : f.size] = f
        K = np.ascontiguousarray(bk.K, dtype=np.int32)
        arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in (bk.ix, bk.iy, bk.sx, bk.sy, bk.spx, bk.spy, bk.ntotx, bk.ntoty)]
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.pfbs_psi_create(self.prec, self.device, self.nband, self.nx, self.ny, self.nbasis, self.nlevel,
                                             _vp(K), _vp(filt), *[_vp(a) for a in arrs], self.nxmax, self.nymax,
                                             C.byref(self._h)))

    # shapes -------------------------------------------------------------------------------
    @property
    def coeff_shape(self):
        tail = (self.nymax, self.nxmax) if self._transposed else (self.nxmax, self.nymax)
        return (self.nband, self.nbasis) + tail

    @property
    def _flags(self):
        return TRANSPOSED if self._transposed else 0

    def _check(self, x, alpha):
        if x.shape != (self.nband, self.nx, self.ny):
            raise ValueError(f"image cube shape {x.shape} != {(self.nband, self.nx, self.ny)}")
        if alpha.shape != self.coeff_shape:
            raise ValueError(f"coefficient cube shape {alpha.shape} != {self.coeff_shape}")

    # reference-facing, numpy in place -----------------------------------------------------
    def dot(self, x, alphao):
        """image to coeffs (alphao is overwritten)."""
        self._check(x, alphao)
        xin = np.ascontiguousarray(x, dtype=self.rdt)
        out = alphao if (alphao.flags.c_contiguous and alphao.dtype == self.rdt) else np.empty(self.coeff_shape, self.rdt)
        _lib.check(self._lib.pfbs_psi_dot(self._h, _vp(xin), _vp(out), _lib.HOST_PTRS | self._flags, None))
        if out is not alphao:
            alphao[...] = out
        return alphao

    def hdot(self, alpha, xo):
        """coeffs to image (xo is overwritten)."""
        self._check(xo, alpha)
        ain = np.ascontiguousarray(alpha, dtype=self.rdt)
        out = xo if (xo.flags.c_contiguous and xo.dtype == self.rdt) else np.empty((self.nband, self.nx, self.ny), self.rdt)
        _lib.check(self._lib.pfbs_psi_hdot(self._h, _vp(ain), _vp(out), _lib.HOST_PTRS | self._flags, None))
        if out is not xo:
            xo[...] = out
        return xo

    # device-resident (torch tensors as buffers; always the x-first layout) -----------------
    def dot_dev(self, x_t, alpha_t, stream=None):
        _lib.check(self._lib.pfbs_psi_dot(self._h, C.c_void_p(x_t.data_ptr()), C.c_void_p(alpha_t.data_ptr()),
                                          _lib.DEVICE_PTRS, stream))

    def hdot_dev(self, alpha_t, x_t, stream=None):
        _lib.check(self._lib.pfbs_psi_hdot(self._h, C.c_void_p(alpha_t.data_ptr()), C.c_void_p(x_t.data_ptr()),
                                           _lib.DEVICE_PTRS, stream))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pfbs_psi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Psi(PsiNocopyt):
    """`Psi` of ``operators/psi.py:549-607``: same dictionary, coefficient cubes ``(nband, nbasis, nymax, nxmax)``."""

    _transposed = True


def _torch():
    import torch

    return torch


class L21:
    """Weighted l21 regulariser over a dictionary (``prox/l21.py:17-91``); the maths runs on the device."""

    def __init__(self, psi, bases=None, nu: float = 1.0, rmsfactor: float = 1.0, alpha: float = 2.0):
        for attr in ("dot", "hdot", "nband", "nbasis", "nxmax", "nymax"):
            if not hasattr(psi, attr):
                raise TypeError(f"psi lacks {attr!r} (PsiOperator protocol)")
        self.psi = psi
        self.nu = nu
        self.bases = tuple(bases) if bases is not None else tuple(getattr(psi, "bases", ()))
        self.rmsfactor, self.alpha = rmsfactor, alpha
        self.l1weight = np.ones(psi.coeff_shape[1:])

    def prox(self, v, vout, lam, sigma=1.0):
        """vout = prox_{(lam/sigma) ||W .||_21}(v / sigma), in place (numpy in / out)."""
        torch = _torch()
        psi = self.psi
        dev = torch.device("cuda", psi.device)
        vt = torch.from_numpy(np.ascontiguousarray(v, dtype=psi.rdt)).to(dev)
        wt = torch.from_numpy(np.ascontiguousarray(self.l1weight, dtype=psi.rdt)).to(dev)
        rt = torch.empty_like(vt)
        nband, ncoef = v.shape[0], int(np.prod(v.shape[1:]))
        s = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(psi._lib.pfbs_prox_21m(psi.prec, psi.device, C.c_void_p(vt.data_ptr()), C.c_void_p(rt.data_ptr()),
                                          C.c_void_p(wt.data_ptr()), float(lam), float(sigma), nband, ncoef, s))
        vout[...] = rt.cpu().numpy()
        return vout

    def prox_dev(self, v_t, out_t, w_t, lam, sigma=1.0, reduce=None):
        """Device tensors; with `reduce` the band sum is completed across ranks before thresholding."""
        psi = self.psi
        torch = _torch()
        nband, ncoef = v_t.shape[0], int(v_t[0].numel())
        s = torch.cuda.current_stream(v_t.device).cuda_stream
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        if reduce is None:
            _lib.check(psi._lib.pfbs_prox_21m(psi.prec, psi.device, p(v_t), p(out_t), p(w_t), float(lam), float(sigma),
                                              nband, ncoef, s))
            return
        # sharded bands: ratio from the global band sum (one all-reduce), then a local scale
        tot = v_t.sum(dim=0)
        reduce(tot)
        tot /= sigma
        a = tot.abs()
        ratio = torch.where(tot != 0, torch.clamp(a - lam * w_t / sigma, min=0.0) / torch.where(a > 0, a, torch.ones_like(a)) / sigma,
                            torch.zeros_like(a))
        torch.mul(v_t, ratio.unsqueeze(0), out=out_t)

    # l1 reweighting (prox/l21.py:52-91, utils/misc.py:750-764): once per major cycle, host side
    @property
    def reweight_active(self) -> bool:
        return getattr(self, "_rms_comps", None) is not None

    def _band_sum_coeffs(self, x, reduce=None):
        psi = self.psi
        out = np.empty(psi.coeff_shape, psi.rdt)
        psi.dot(np.ascontiguousarray(x, dtype=psi.rdt), out)
        tot = out.sum(axis=0)
        return reduce(tot) if reduce is not None else tot

    def init_reweighting(self, update, reduce=None):
        tmp = self._band_sum_coeffs(update, reduce)
        rms = np.ones(self.psi.nbasis)
        for i in range(self.psi.nbasis):
            nz = tmp[i][tmp[i] != 0]
            if nz.size:
                rms[i] = np.std(nz)
        self._rms_comps = rms

    def update_weights(self, x, reduce=None):
        if not self.reweight_active:
            raise RuntimeError("init_reweighting() has not been called")
        mcomps = np.abs(self._band_sum_coeffs(x, reduce))
        r = self._rms_comps[:, None, None]
        self.l1weight = (1 + self.rmsfactor) / (1 + mcomps ** self.alpha / r ** self.alpha)
        return self.l1weight

    def dual_update(self, vp, v, lam, sigma=1.0):
        """Fused dual update of ``dual_update_numba_fast`` (v updated in place, numpy in / out)."""
        torch = _torch()
        psi = self.psi
        dev = torch.device("cuda", psi.device)
        vpt = torch.from_numpy(np.ascontiguousarray(vp, dtype=psi.rdt)).to(dev)
        vt = torch.from_numpy(np.ascontiguousarray(v, dtype=psi.rdt)).to(dev)
        wt = torch.from_numpy(np.ascontiguousarray(self.l1weight, dtype=psi.rdt)).to(dev)
        self.dual_update_dev(vpt, vt, wt, lam, sigma)
        v[...] = vt.cpu().numpy()
        return v

    def dual_update_dev(self, vp_t, v_t, w_t, lam, sigma, reduce=None, bsum_t=None, vbar_t=None):
        """Device tensors.  With `reduce` (an in-place all-reduce of a device tensor over the ranks that hold
        the other bands) the update runs as: local band sum -> reduce -> scale.  `vbar_t` (optional) receives the
        extrapolated dual 2 v - vp in the same pass."""
        psi = self.psi
        torch = _torch()
        nband, ncoef = v_t.shape[0], int(v_t[0].numel())
        s = torch.cuda.current_stream(v_t.device).cuda_stream
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        if reduce is None:
            _lib.check(psi._lib.pfbs_dual_update(psi.prec, psi.device, p(vp_t), p(v_t), p(w_t), float(lam), float(sigma),
                                                 nband, ncoef, None, 0, p(vbar_t), s))
            return
        if bsum_t is None:
            bsum_t = torch.empty_like(v_t[0])
        _lib.check(psi._lib.pfbs_dual_update(psi.prec, psi.device, p(vp_t), p(v_t), p(w_t), float(lam), float(sigma),
                                             nband, ncoef, p(bsum_t), 1, None, s))
        reduce(bsum_t)
        _lib.check(psi._lib.pfbs_dual_update(psi.prec, psi.device, p(vp_t), p(v_t), p(w_t), float(lam), float(sigma),
                                             nband, ncoef, p(bsum_t), 2, p(vbar_t), s))


class PrimalDual:
    """Primal-dual solver for ``min_x f(x) + lam * g(Psi^T x)`` (``opt/primal_dual.py:284-448``).

    `set_grad(grad)` takes either the reference's numpy callable ``grad(xp) -> (nband, nx, ny)`` (the cube then
    makes one host round trip per iteration) or an object with ``grad.device_apply(x_t, out_t)`` working on
    device tensors (e.g. a PSF-convolution Hessian minus the dirty image), in which case nothing leaves the
    device until the loop ends.  `reduce_tensor` / `reduce_scalars` are the cross-band all-reduces for
    band-sharded cubes (dist.py); None = every band is local.
    """

    def __init__(self, tol=1e-5, maxit=1000, report_freq=10, verbosity=1, gamma=1.0, sigma=None, on_converge=None,
                 positivity=0, reduce_tensor=None, reduce_scalars=None):
        self.tol, self.maxit, self.report_freq, self.verbosity = tol, maxit, report_freq, verbosity
        self.gamma, self._sigma_opt, self.on_converge = gamma, sigma, on_converge
        self.positivity = int(positivity)
        self.reduce_tensor, self.reduce_scalars = reduce_tensor, reduce_scalars
        self._grad = None
        self._reg = None
        self._v = None
        self._work = None
        self.niter = 0
        self.eps = 1.0

    def setup(self, prox, hessnorm: float) -> None:
        self._reg = prox
        self._work = None
        self.hessnorm = hessnorm
        nu = prox.nu
        sigma = self._sigma_opt
        if sigma is None:
            sigma = hessnorm / (2.0 * self.gamma) / nu
        self.sigma = sigma
        self.tau = 0.98 / (hessnorm / (2.0 * self.gamma) + sigma * nu ** 2)
        psi = prox.psi
        torch = _torch()
        self._dev = torch.device("cuda", psi.device)
        self._tdt = torch.float32 if psi.rdt == np.float32 else torch.float64
        # the dual lives on the device in the x-first layout, warm-started across solve() calls
        self._v = torch.zeros((psi.nband, psi.nbasis, psi.nxmax, psi.nymax), dtype=self._tdt, device=self._dev)

    def set_grad(self, grad) -> None:
        self._grad = grad

    def reset(self) -> None:
        if self._v is not None:
            self._v.zero_()

    @property
    def dual(self):
        """Host copy of the dual in the layout of the bound dictionary."""
        v = self._v.cpu().numpy()
        return np.ascontiguousarray(v.transpose(0, 1, 3, 2)) if self._reg.psi._transposed else v

    def solve(self, x, lam: float):
        if self._reg is None:
            raise RuntimeError("regulariser not bound; call setup() before solve()")
        if self._grad is None:
            raise RuntimeError("grad not set; call set_grad() before solve()")
        torch = _torch()
        reg, psi = self._reg, self._reg.psi
        lib, prec, dv = psi._lib, psi.prec, psi.device
        # work buffers live with the solver: solve() is called once per major cycle and the cubes are GB-sized
        wk = self._work
        if wk is None or wk["x"].shape != tuple(np.shape(x)):
            shp = tuple(np.shape(x))
            wk = self._work = dict(
                x=torch.empty(shp, dtype=self._tdt, device=self._dev), xp=torch.empty(shp, dtype=self._tdt, device=self._dev),
                xout=torch.empty(shp, dtype=self._tdt, device=self._dev), g=torch.empty(shp, dtype=self._tdt, device=self._dev),
                v=torch.empty_like(self._v), vbar=torch.empty_like(self._v),
                pin=torch.empty(shp, dtype=self._tdt, pin_memory=True), w=None, w_src=None)
        x_t, xp_t, xout_t, g_t = wk["x"], wk["xp"], wk["xout"], wk["g"]
        pin_np = wk["pin"].numpy()
        np.copyto(pin_np, x, casting="same_kind")
        x_t.copy_(wk["pin"], non_blocking=True)
        xp_t.copy_(x_t)
        # three dual buffers: vp (previous iterate), v (receives Psi^T xp, becomes the new iterate in place) and
        # vbar (extrapolated 2 v - vp, written by the fused dual update); vp and v swap roles every iteration
        vp_t = self._v
        v_t = wk["v"]
        vbar_t = wk["vbar"]
        if wk["w_src"] is not reg.l1weight:  # the regulariser replaces the array when it reweights (prox/l21.py:88-91)
            w = reg.l1weight.transpose(0, 2, 1) if psi._transposed else reg.l1weight
            wk["w"] = torch.from_numpy(np.ascontiguousarray(w, dtype=psi.rdt)).to(self._dev)
            wk["w_src"] = reg.l1weight
        w_t = wk["w"]
        bsum_t = torch.empty_like(v_t[0]) if self.reduce_tensor is not None else None
        s = torch.cuda.current_stream(self._dev).cuda_stream
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        npix = psi.nx * psi.ny
        dev_grad = getattr(self._grad, "device_apply", None)
        nd = (C.c_double * 2)()
        eps, k = 1.0, 0
        for k in range(self.maxit):
            psi.dot_dev(xp_t, v_t, s)
            reg.dual_update_dev(vp_t, v_t, w_t, lam, self.sigma, self.reduce_tensor, bsum_t, vbar_t)
            psi.hdot_dev(vbar_t, xout_t, s)
            if dev_grad is not None:
                dev_grad(xp_t, g_t)
            else:
                g_t.copy_(torch.from_numpy(np.ascontiguousarray(self._grad(xp_t.cpu().numpy()), dtype=psi.rdt)))
            xout_t += g_t
            pos = self.positivity
            if pos == 2 and self.reduce_tensor is not None:
                pos = 0  # needs the minimum over ALL bands: done below with one all-reduce
            _lib.check(lib.pfbs_primal_step(prec, dv, p(x_t), p(xp_t), p(xout_t), float(self.tau), pos, psi.nband, npix, s))
            if self.positivity == 2 and pos == 0:
                neg = (x_t <= 0).any(dim=0).to(x_t.dtype)
                self.reduce_tensor(neg)
                x_t.mul_((neg == 0).to(x_t.dtype))
            _lib.check(lib.pfbs_norm_diff(prec, dv, p(x_t), p(xp_t), x_t.numel(), nd, s))
            num, den = float(nd[0]), float(nd[1])
            if self.reduce_scalars is not None:
                num, den = (float(t) for t in self.reduce_scalars(np.array([num, den])))
            eps = float(np.sqrt(num / max(den, 1e-12))) if den > 0 else 1.0
            if eps < self.tol:
                if self.on_converge is None or self.on_converge(x_t.cpu().numpy(), k, eps):
                    break
            xp_t.copy_(x_t)
            vp_t, v_t = v_t, vp_t  # the new iterate becomes the previous one; its buffer is overwritten next
            if not k % self.report_freq and self.verbosity > 1:
                print(f"PD: at iteration {k} eps = {eps:.3e}")
        else:
            vp_t, v_t = v_t, vp_t  # loop ran out after a swap: the latest iterate sits in vp_t
        self._v = v_t  # warm start of the next solve (on break: v_t holds the latest iterate)
        wk["v"] = vp_t  # the other dual buffer is free for the next call
        self.niter, self.eps = k, eps
        if self.verbosity:
            print(f"PD: max iters reached, eps = {eps:.3e}" if k == self.maxit - 1 else f"PD: converged after {k} iterations")
        wk["pin"].copy_(x_t)  # device -> pinned host (synchronous), then one host copy into the caller's array
        if isinstance(x, np.ndarray) and x.shape == pin_np.shape and x.flags.writeable:
            np.copyto(x, pin_np, casting="same_kind")
            return x
        return pin_np.astype(psi.rdt, copy=True)


class ForwardBackward:
    """Forward-backward splitting with optional FISTA momentum (``opt/forward_backward.py:21-133``), the cube
    resident on the device.  Same `setup` / `set_grad` / `solve` contract and grad conventions as `PrimalDual`."""

    def __init__(self, tol=1e-5, maxit=1000, report_freq=10, verbosity=1, gamma=1.0, acceleration=True, on_converge=None,
                 positivity=0, reduce_tensor=None, reduce_scalars=None):
        self.tol, self.maxit, self.report_freq, self.verbosity = tol, maxit, report_freq, verbosity
        self.gamma, self.acceleration, self.on_converge = gamma, acceleration, on_converge
        self.positivity = int(positivity)
        self.reduce_tensor, self.reduce_scalars = reduce_tensor, reduce_scalars
        self._grad = self._reg = None
        self.niter, self.eps = 0, 1.0

    def setup(self, prox, hessnorm: float) -> None:
        self._reg = prox
        self.hessnorm = hessnorm
        self.step = 2.0 * self.gamma / hessnorm

    def set_grad(self, grad) -> None:
        self._grad = grad

    def reset(self) -> None:
        """No warm-start state beyond x itself."""

    def solve(self, x, lam: float):
        if self._reg is None:
            raise RuntimeError("regulariser not bound; call setup() before solve()")
        if self._grad is None:
            raise RuntimeError("grad not set; call set_grad() before solve()")
        torch = _torch()
        reg, psi = self._reg, self._reg.psi
        lib, prec, dv = psi._lib, psi.prec, psi.device
        dev = torch.device("cuda", dv)
        tdt = torch.float32 if psi.rdt == np.float32 else torch.float64
        x_t = torch.from_numpy(np.ascontiguousarray(x, dtype=psi.rdt)).to(dev)
        xp_t, y_t = x_t.clone(), x_t.clone()
        g_t, xout_t = torch.empty_like(x_t), torch.empty_like(x_t)
        a_t = torch.empty((psi.nband, psi.nbasis, psi.nxmax, psi.nymax), dtype=tdt, device=dev)
        b_t = torch.empty_like(a_t)
        w = reg.l1weight.transpose(0, 2, 1) if psi._transposed else reg.l1weight
        w_t = torch.from_numpy(np.ascontiguousarray(w, dtype=psi.rdt)).to(dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        ax = lambda out, a, xx, b, yy: _lib.check(lib.pfbs_axpby(prec, dv, p(out), float(a), p(xx), float(b), p(yy), out.numel(), s))  # noqa: E731
        dev_grad = getattr(self._grad, "device_apply", None)
        nd = (C.c_double * 2)()
        t, eps, k = 1.0, 1.0, 0
        for k in range(self.maxit):
            if dev_grad is not None:
                dev_grad(y_t, g_t)
            else:
                g_t.copy_(torch.from_numpy(np.ascontiguousarray(self._grad(y_t.cpu().numpy()), dtype=psi.rdt)))
            ax(x_t, 1.0, y_t, -self.step, g_t)
            # tight-frame prox: x += Psi (prox(Psi^T x) - Psi^T x) / nu
            psi.dot_dev(x_t, a_t, s)
            reg.prox_dev(a_t, b_t, w_t, self.step * lam, 1.0, self.reduce_tensor)
            ax(b_t, 1.0, b_t, -1.0, a_t)
            psi.hdot_dev(b_t, xout_t, s)
            ax(x_t, 1.0, x_t, 1.0 / reg.nu, xout_t)
            if self.positivity:
                pos = self.positivity
                if pos == 2 and self.reduce_tensor is not None:
                    neg = (x_t <= 0).any(dim=0).to(x_t.dtype)
                    self.reduce_tensor(neg)
                    x_t.mul_((neg == 0).to(x_t.dtype))
                else:
                    _lib.check(lib.pfbs_primal_step(prec, dv, p(x_t), p(x_t), p(x_t), 0.0, pos, psi.nband, psi.nx * psi.ny, s))
            _lib.check(lib.pfbs_norm_diff(prec, dv, p(x_t), p(xp_t), x_t.numel(), nd, s))
            num, den = float(nd[0]), float(nd[1])
            if self.reduce_scalars is not None:
                num, den = (float(v) for v in self.reduce_scalars(np.array([num, den])))
            eps = float(np.sqrt(num / max(den, 1e-12))) if den > 0 else 1.0
            if eps < self.tol:
                if self.on_converge is None or self.on_converge(x_t.cpu().numpy(), k, eps):
                    break
            if self.acceleration:
                tp = t
                t = (1.0 + np.sqrt(1.0 + 4.0 * tp ** 2)) / 2.0
                c = (tp - 1.0) / t
                ax(y_t, 1.0 + c, x_t, -c, xp_t)
            else:
                y_t.copy_(x_t)
            xp_t.copy_(x_t)
            if not k % self.report_freq and self.verbosity > 1:
                print(f"FB: at iteration {k} eps = {eps:.3e}")
        self.niter, self.eps = k, eps
        if self.verbosity:
            print(f"FB: max iters reached, eps = {eps:.3e}" if k == self.maxit - 1 else f"FB: converged after {k} iterations")
        return x_t.cpu().numpy()
