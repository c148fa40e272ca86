"""GPU: device PSF-convolution Hessians against the numpy restatement of
/root/reference/src/pfb_imaging/operators/psf.py:8-31 and operators/hessian.py:103-143 (r2c, multiply, c2r, crop)."""
import numpy as np
import pytest

from pfb_imaging_b200 import psf as P
from pfbg_testutil import rel_l2

pytestmark = pytest.mark.gpu


def ref_hessian_psf(x, abspsf, ny_psf, beam=None, eta=None):
    nx, ny = x.shape
    xpad = np.zeros((abspsf.shape[0], ny_psf))
    xpad[:nx, :ny] = x if beam is None else x * beam
    xhat = np.fft.rfft2(xpad) * abspsf
    out = np.fft.irfft2(xhat, s=xpad.shape)[:nx, :ny].copy()
    if beam is not None:
        out *= beam
    if eta:
        out += x * eta
    return out


@pytest.mark.parametrize("nx,ny,nxp,nyp", [(64, 48, 96, 80), (100, 100, 140, 154), (128, 128, 176, 176), (32, 32, 32, 32),
                                           (51, 37, 72, 60), (33, 64, 64, 90)])  # odd row counts: last row pair is half empty
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_hessian_psf_slice_matches_numpy(gpu, nx, ny, nxp, nyp, dt):
    rng = np.random.default_rng(nx + nyp)
    psf = np.zeros((nxp, nyp))
    psf[nxp // 2 - 8: nxp // 2 + 8, nyp // 2 - 8: nyp // 2 + 8] = rng.standard_normal((16, 16))
    psf[nxp // 2, nyp // 2] = 10.0
    psfhat = np.fft.rfft2(np.fft.ifftshift(psf))
    abspsf = np.abs(psfhat)
    x = rng.standard_normal((nx, ny))
    beam = rng.uniform(0.5, 1.0, (nx, ny))
    tol = 1e-12 if dt == np.float64 else 2e-5
    got = P.hessian_psf_slice(x.astype(dt), abspsf=abspsf.astype(dt), beam=beam.astype(dt), lastsize=nyp, eta=0.3)
    assert got.dtype == dt
    assert rel_l2(got, ref_hessian_psf(x, abspsf, nyp, beam, 0.3)) <= tol
    # complex psfhat through psf_convolve_slice (operators/psf.py:8-31)
    xout = np.empty((nx, ny), dtype=dt)
    P.psf_convolve_slice(None, None, xout, psfhat.astype(np.complex64 if dt == np.float32 else np.complex128), nyp, x.astype(dt))
    ref = np.fft.irfft2(np.fft.rfft2(np.pad(x, ((0, nxp - nx), (0, nyp - ny)))) * psfhat, s=(nxp, nyp))[:nx, :ny]
    assert rel_l2(xout, ref) <= tol
    P.clear_convolver_cache()


def test_hesspsf_cube_dot_and_idot(gpu):
    rng = np.random.default_rng(3)
    nband, nx, ny, nxp, nyp = 2, 48, 48, 72, 72
    abspsf = np.empty((nband, nxp, nyp // 2 + 1))
    for b in range(nband):
        psf = np.zeros((nxp, nyp))
        psf[nxp // 2 - 3: nxp // 2 + 4, nyp // 2 - 3: nyp // 2 + 4] = np.outer(np.hanning(7), np.hanning(7)) * (1 + b)
        abspsf[b] = np.abs(np.fft.rfft2(np.fft.ifftshift(psf)))
    beam = rng.uniform(0.6, 1.0, (nband, nx, ny))
    H = P.HessPSF(nx, ny, abspsf, beam=beam, eta=0.5, cgtol=1e-9, cgmaxit=500, cgverbose=0)
    x = rng.standard_normal((nband, nx, ny))
    hx = H.dot(x)
    for b in range(nband):
        assert rel_l2(hx[b], ref_hessian_psf(x[b], abspsf[b], nyp, beam[b], 0.5)) <= 1e-12
    xr = H.idot(hx, mode="psf")
    assert rel_l2(xr, x) <= 1e-5
    with pytest.raises(ValueError):  # hessian.py:436 ("Unknown mode"); "direct" is covered in test_gpu_dropins.py
        H.idot(hx, mode="cheap")
    H.close()
