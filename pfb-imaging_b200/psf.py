"""PSF-convolution Hessians on the device: the operator pfb-imaging iterates inside
pcg, the power method and the primal-dual loop.  Drop-ins for (/root/reference/src/pfb_imaging):

  psf_convolve_slice   operators/psf.py:8-31
  hessian_psf_slice    operators/hessian.py:103-143
  HessPSF.dot / idot(mode="psf")   operators/hessian.py:251-436

numpy in, numpy out; the transforms run in ``libpfbgrid.so`` (``csrc/psfconv.cuh``: row pass, one
transform-multiply-transform column kernel that never leaves shared memory, row pass back).  The
``xpad`` / ``xhat`` scratch arguments of the reference signatures are accepted and ignored.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .wgridder import current_device

_RDT = {np.dtype(np.float32): (_lib.PFBG_F32, np.complex64), np.dtype(np.float64): (_lib.PFBG_F64, np.complex128)}


class PsfConvolver:
    """out = beam * crop(IFFT(FFT(pad(beam * x)) * khat)) + eta * x for one band."""

    def __init__(self, nx, ny, nx_psf, ny_psf, dtype=np.float64, device=None):
        self.rdt = np.dtype(dtype)
        if self.rdt not in _RDT:
            raise TypeError("dtype must be float32 or float64")
        self.prec, self.cdt = _RDT[self.rdt]
        self.nx, self.ny, self.nx_psf, self.ny_psf = int(nx), int(ny), int(nx_psf), int(ny_psf)
        self.device = current_device() if device is None else int(device)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.pfbg_conv_create(self.prec, self.device, self.nx, self.ny, self.nx_psf, self.ny_psf,
                                              C.byref(self._h)))

    def set_kernel(self, khat):
        """khat: (nx_psf, ny_psf//2+1) r2c half spectrum (real ``abspsf`` or complex ``psfhat``) or the
        full (nx_psf, ny_psf) spectrum."""
        khat = np.asarray(khat)
        if khat.shape == (self.nx_psf, self.ny_psf // 2 + 1):
            half = 1
        elif khat.shape == (self.nx_psf, self.ny_psf):
            half = 0
        else:
            raise ValueError(f"kernel shape {khat.shape} matches neither the half nor the full spectrum")
        k = np.ascontiguousarray(khat, dtype=self.cdt)
        _lib.check(self._lib.pfbg_conv_set_kernel(self._h, C.c_void_p(k.ctypes.data), half, _lib.HOST_PTRS, None))
        return self

    def apply(self, x, beam=None, eta=None, out=None):
        x = np.ascontiguousarray(x, dtype=self.rdt)
        if x.shape != (self.nx, self.ny):
            raise ValueError(f"x shape {x.shape} != ({self.nx}, {self.ny})")
        if beam is not None:
            beam = np.ascontiguousarray(beam, dtype=self.rdt)
            if beam.shape != x.shape:
                raise ValueError("beam shape does not match the image")
        res = np.empty_like(x) if out is None else out
        tmp = res if (res.flags.c_contiguous and res.dtype == self.rdt) else np.empty_like(x)
        _lib.check(self._lib.pfbg_conv_apply(self._h, C.c_void_p(x.ctypes.data),
                                             None if beam is None else C.c_void_p(beam.ctypes.data),
                                             float(eta) if eta else 0.0, C.c_void_p(tmp.ctypes.data),
                                             _lib.HOST_PTRS, None))
        if tmp is not res:
            res[...] = tmp
        return res

    def apply_dev(self, x_ptr, out_ptr, beam_ptr=None, eta=None, stream=None):
        """Same operator on device pointers (asynchronous on `stream`): what the device-resident deconvolution loops call."""
        _lib.check(self._lib.pfbg_conv_apply(self._h, C.c_void_p(int(x_ptr)), None if beam_ptr is None else C.c_void_p(int(beam_ptr)),
                                             float(eta) if eta else 0.0, C.c_void_p(int(out_ptr)), _lib.DEVICE_PTRS, stream))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pfbg_conv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CONV_CACHE: dict = {}


def _convolver_for(khat, nx, ny, lastsize, rdt):
    khat = np.asarray(khat)
    nx_psf = khat.shape[0]
    ny_psf = int(lastsize) if lastsize is not None else 2 * (khat.shape[1] - 1)
    key = (khat.ctypes.data, khat.shape, khat.dtype.str, nx, ny, ny_psf, np.dtype(rdt).str)
    chk = float(np.abs(khat.ravel()[:: max(1, khat.size // 1024)]).sum())
    hit = _CONV_CACHE.get(key)
    if hit is not None and hit[1] == chk:
        return hit[0]
    if hit is not None:
        hit[0].close()
    while len(_CONV_CACHE) >= 16:
        _CONV_CACHE.pop(next(iter(_CONV_CACHE)))[0].close()
    cv = PsfConvolver(nx, ny, nx_psf, ny_psf, dtype=rdt).set_kernel(khat)
    _CONV_CACHE[key] = (cv, chk, khat)
    return cv


def clear_convolver_cache():
    while _CONV_CACHE:
        _CONV_CACHE.popitem()[1][0].close()


def psf_convolve_slice(xpad, xhat, xout, psfhat, lastsize, x, nthreads=1):
    """operators/psf.py:8-31."""
    x = np.asarray(x)
    cv = _convolver_for(psfhat, x.shape[0], x.shape[1], lastsize, x.dtype)
    xout[...] = cv.apply(x)
    return xout


def hessian_psf_slice(x, xpad=None, xhat=None, xout=None, abspsf=None, beam=None, lastsize=None, nthreads=1, eta=None):
    """Tikhonov regularised PSF-convolution Hessian (operators/hessian.py:103-143)."""
    x = np.asarray(x)
    cv = _convolver_for(abspsf, x.shape[0], x.shape[1], lastsize, x.dtype)
    res = cv.apply(x, beam=beam, eta=eta)
    if xout is None:
        return res
    np.copyto(xout, res)
    return xout


class HessPSF:
    """Cube-level PSF-convolution Hessian with the reference's constructor and ``dot`` / ``idot(mode="psf")``
    (operators/hessian.py:251-436); one device convolver per band."""

    def __init__(self, nx, ny, abspsf, beam=None, eta=1.0, nthreads=1, cgtol=1e-3, cgmaxit=300, cgverbose=2, cgrf=25,
                 taper_width=32, min_beam=5e-3):
        self.nx, self.ny = nx, ny
        self.abspsf = abspsf
        self.nband, self.nx_psf, self.nyo2 = abspsf.shape
        self.ny_psf = 2 * (self.nyo2 - 1)
        if beam is not None and not (np.asarray(beam) == 1).all():
            if beam.shape != (self.nband, nx, ny):
                raise AssertionError("beam must have shape (nband, nx, ny)")
            self.beam = beam
        else:
            self.beam = (None,) * self.nband
        self.eta = np.tile(eta, self.nband) if isinstance(eta, float) else np.array(eta)
        if self.eta.size != self.nband:
            raise AssertionError("eta must be a float or have one entry per band")
        self.cgtol, self.cgmaxit, self.cgverbose, self.cgrf = cgtol, cgmaxit, cgverbose, cgrf
        self.conv = [PsfConvolver(nx, ny, self.nx_psf, self.ny_psf, dtype=np.float64).set_kernel(abspsf[b])
                     for b in range(self.nband)]

    def set_beam(self, beam):
        assert beam.shape == (self.nband, self.nx, self.ny)
        self.beam = beam

    def _cube(self, x):
        if x.ndim == 3:
            return x
        if x.ndim == 2:
            return x[None]
        raise ValueError("Unsupported number of input dimensions")

    def dot(self, x):
        xt = self._cube(x)
        if xt.shape != (self.nband, self.nx, self.ny):
            raise AssertionError("input shape does not match the operator")
        out = np.empty((self.nband, self.nx, self.ny))
        for b in range(self.nband):
            self.conv[b].apply(xt[b], beam=self.beam[b], eta=float(self.eta[b]), out=out[b])
        return out

    hdot = dot

    def idot(self, x, mode="psf", x0=None):
        """Per-band CG inverse of `dot` (hessian.py:408-432, mode="psf")."""
        if mode != "psf":
            raise ValueError(f"mode {mode!r} is not provided by the device operator (only 'psf')")
        from .solvers import pcg

        xt = self._cube(x)
        x0 = np.zeros_like(xt) if x0 is None else self._cube(x0)
        out = np.empty((self.nband, self.nx, self.ny))
        for b in range(self.nband):
            hess = lambda v, b=b: self.conv[b].apply(v, beam=self.beam[b], eta=float(self.eta[b]))
            out[b] = pcg(hess, np.ascontiguousarray(xt[b], dtype=np.float64), x0=np.array(x0[b], dtype=np.float64),
                         tol=self.cgtol, maxit=self.cgmaxit, minit=3, verbosity=min(self.cgverbose, 1) if self.cgverbose < 2 else 0,
                         report_freq=self.cgrf, backtrack=False, return_resid=False)
        return out

    def close(self):
        for c in self.conv:
            c.close()


class PsfGradient:
    """Gradient of the smooth term of the deconvolution sub-problem, ``grad(x) = hess(x) - dirty`` with the
    PSF-convolution Hessian (the `grad` callables built in ``core/sara.py:300-340`` / ``deconv`` presets), usable both as
    the reference's numpy callable and, through ``device_apply``, inside the device-resident loops of
    ``pfb_imaging_b200.sara`` (no host round trip per iteration)."""

    def __init__(self, hess: HessPSF, dirty):
        self.hess = hess
        self.dirty = np.ascontiguousarray(dirty, dtype=np.float64)
        if self.dirty.shape != (hess.nband, hess.nx, hess.ny):
            raise ValueError("dirty must have shape (nband, nx, ny)")
        self._dev = None

    def __call__(self, x):
        return self.hess.dot(x) - self.dirty

    def device_apply(self, x_t, out_t):
        import torch

        if self._dev is None:
            dev = x_t.device
            self._dev = dict(dirty=torch.from_numpy(self.dirty).to(dev),
                             beam=[None if b is None else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).to(dev)
                                   for b in self.hess.beam])
        s = torch.cuda.current_stream(x_t.device).cuda_stream
        for b in range(self.hess.nband):
            bm = self._dev["beam"][b]
            self.hess.conv[b].apply_dev(x_t[b].data_ptr(), out_t[b].data_ptr(), None if bm is None else bm.data_ptr(),
                                        float(self.hess.eta[b]), s)
        out_t -= self._dev["dirty"]
