"""GPU: the per-partition products / residual properties of /root/reference/tests/test_imager_pass2.py
(row additivity, robust re-weighting bound, zero model, partition additivity, beam applied once),
on the same synthetic partitions, with a stand-in for xarray (absent here)."""
import numpy as np
import pytest

from pfb_imaging_b200 import operators as ops
from pfb_imaging_b200.weighting import _compute_counts

pytestmark = pytest.mark.gpu


class V:
    def __init__(self, v):
        self.values = np.asarray(v)


class Part:
    def __init__(self, **kw):
        self.attrs = kw.pop("attrs", {})
        for k, v in kw.items():
            setattr(self, k, V(v))


def _synth_partition(nrow=200, seed=0):
    rng = np.random.default_rng(seed)
    uvw = rng.standard_normal((nrow, 3)) * 100.0
    freq = np.array([1.0e9])
    vis = rng.standard_normal((1, nrow, 1)) + 1j * rng.standard_normal((1, nrow, 1))
    wgt = np.abs(rng.standard_normal((1, nrow, 1))) + 0.1
    mask = np.ones((nrow, 1), dtype=np.uint8)
    return Part(VIS=vis, WEIGHT=wgt, MASK=mask, UVW=uvw, FREQ=freq, BEAM=np.ones((1, 3, 3)),
                l_beam=np.array([-1.0, 0.0, 1.0]), m_beam=np.array([-1.0, 0.0, 1.0]))


def _cat(p0, p1):
    return Part(VIS=np.concatenate([p0.VIS.values, p1.VIS.values], axis=1),
                WEIGHT=np.concatenate([p0.WEIGHT.values, p1.WEIGHT.values], axis=1),
                MASK=np.concatenate([p0.MASK.values, p1.MASK.values], axis=0),
                UVW=np.concatenate([p0.UVW.values, p1.UVW.values], axis=0), FREQ=p0.FREQ.values,
                BEAM=p0.BEAM.values, l_beam=p0.l_beam.values, m_beam=p0.m_beam.values)


KW = dict(nx=16, ny=16, nx_psf=32, ny_psf=32, cell_rad=1.0e-6, robustness=None)


def test_grid_partition_shapes_and_wsum(gpu):
    part = _synth_partition()
    out = ops.grid_partition(part, None, **KW)
    assert out["DIRTY"].shape == (1, 16, 16) and out["PSF"].shape == (1, 32, 32)
    assert out["PSFHAT"].shape == (1, 32, 32 // 2 + 1) and out["BEAM"].shape == (1, 16, 16)
    assert out["WSUM"].shape == (1,)
    expected = (part.WEIGHT.values[0] * part.MASK.values).sum()
    np.testing.assert_allclose(out["WSUM"][0], expected, rtol=1e-6)
    assert np.isfinite(out["DIRTY"]).all()
    # PSF peak = wsum at the centre pixel (unit visibilities)
    np.testing.assert_allclose(out["PSF"][0, 16, 16], out["WSUM"][0], rtol=1e-6)


def test_grid_partition_row_additivity(gpu):
    p0, p1 = _synth_partition(120, 0), _synth_partition(80, 1)
    o_cat, o0, o1 = (ops.grid_partition(p, None, **KW) for p in (_cat(p0, p1), p0, p1))
    np.testing.assert_allclose(o_cat["DIRTY"], o0["DIRTY"] + o1["DIRTY"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(o_cat["PSF"], o0["PSF"] + o1["PSF"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(o_cat["WSUM"], o0["WSUM"] + o1["WSUM"], rtol=1e-6)


def test_grid_partition_robust_reweights(gpu):
    part = _synth_partition()
    nx_pad = ny_pad = 32
    counts = _compute_counts(part.UVW.values, part.FREQ.values, part.MASK.values, part.WEIGHT.values, nx_pad, ny_pad,
                             1.0e-6, 1.0e-6, part.WEIGHT.values.dtype, ngrid=1, usign=-1.0, vsign=1.0)
    out = ops.grid_partition(part, counts, nx=16, ny=16, nx_psf=32, ny_psf=32, cell_rad=1.0e-6, robustness=-2.0,
                             nx_pad=nx_pad, ny_pad=ny_pad)
    assert out["WEIGHT"].shape == part.WEIGHT.values.shape
    assert out["WEIGHT"].max() <= part.WEIGHT.values.max() + 1e-9
    nat = ops.grid_partition(part, None, **KW)
    np.testing.assert_allclose(nat["WEIGHT"], part.WEIGHT.values)


def _image_beam_partition(nx, ny, nrow=200, seed=0, beam_val=1.0, l0=0.0, m0=0.0):
    rng = np.random.default_rng(seed)
    uvw = rng.standard_normal((nrow, 3)) * 100.0
    wgt = np.abs(rng.standard_normal((1, nrow, 1))) + 0.1
    return Part(WEIGHT=wgt, MASK=np.ones((nrow, 1), dtype=np.uint8), UVW=uvw, FREQ=np.array([1.0e9]),
                BEAM=np.full((1, nx, ny), float(beam_val)), attrs={"l0": l0, "m0": m0})


def test_residual_zero_model_returns_dirty(gpu):
    part = _image_beam_partition(16, 16)
    dirty = np.random.default_rng(5).standard_normal((1, 16, 16))
    res = ops.residual_from_partitions(dirty, [part], np.zeros((1, 16, 16)), cell_rad=1.0e-6)
    np.testing.assert_allclose(res, dirty, atol=1e-12)


def test_residual_partition_additivity_and_beam_once(gpu):
    p0, p1 = _image_beam_partition(16, 16, 120, 0), _image_beam_partition(16, 16, 80, 1)
    rng = np.random.default_rng(7)
    dirty, model = rng.standard_normal((1, 16, 16)), rng.standard_normal((1, 16, 16))
    c01 = dirty - ops.residual_from_partitions(dirty, [p0, p1], model, 1.0e-6)
    c0 = dirty - ops.residual_from_partitions(dirty, [p0], model, 1.0e-6)
    c1 = dirty - ops.residual_from_partitions(dirty, [p1], model, 1.0e-6)
    np.testing.assert_allclose(c01, c0 + c1, rtol=1e-5, atol=1e-8)
    pa, pb = _image_beam_partition(16, 16, seed=0, beam_val=1.0), _image_beam_partition(16, 16, seed=0, beam_val=2.0)
    z = np.zeros((1, 16, 16))
    ca = z - ops.residual_from_partitions(z, [pa], model, 1.0e-6)
    cb = z - ops.residual_from_partitions(z, [pb], model, 1.0e-6)
    np.testing.assert_allclose(cb, 2.0 * ca, rtol=1e-5, atol=1e-8)
    ops.clear_plan_cache()


def test_hessian_matches_psf_convolution(gpu):
    """/root/reference/tests/test_hessian_approx.py:234-307: without the w-term the exact Hessian
    equals the (zero-padded FFT) convolution with the PSF."""
    from pfb_imaging_b200 import wgridder as W
    from pfbg_testutil import seed42_array

    _, pix, uvw, freq = seed42_array(nsub=5)
    uvw = 0.02 * uvw
    nx = ny = 64
    nxp = nyp = 128
    cell = pix * 8
    eps = 1e-10
    fu, fv, fw, x0, y0 = ops.wgridder_conventions(0.0, 0.0)
    nrow, nchan = uvw.shape[0], freq.size
    ones = np.ones((nrow, nchan), dtype=np.complex128)
    psf = W.vis2dirty(uvw=uvw, freq=freq, vis=ones, wgt=None, npix_x=nxp, npix_y=nyp, pixsize_x=cell, pixsize_y=cell,
                      center_x=x0, center_y=y0, flip_u=fu, flip_v=fv, flip_w=fw, epsilon=eps, do_wgridding=False,
                      divide_by_n=False)
    psfhat = np.fft.rfft2(np.fft.ifftshift(psf))
    x = np.zeros((nx, ny))
    x[nx // 2, ny // 2] = 1.0
    x[10, 50] = -0.5
    res1 = ops.hessian_slice(x, uvw=uvw, weight=np.ones((nrow, nchan)), vis_mask=np.ones((nrow, nchan), np.uint8),
                             freq=freq, cell=cell, x0=x0, y0=y0, flip_u=fu, flip_v=fv, flip_w=fw, do_wgridding=False,
                             epsilon=eps, double_accum=True)
    xpad = np.zeros((nxp, nyp))
    xpad[:nx, :ny] = x
    conv = np.fft.irfft2(np.fft.rfft2(xpad) * psfhat, s=(nxp, nyp))
    res2 = conv[:nx, :ny]  # psf_convolve_slice (operators/psf.py:8-31): pad at [0:nx,0:ny], crop the same corner
    scale = np.abs(res2).max()
    assert np.allclose(1 + (res2 - res1) / scale, 1)
    ops.clear_plan_cache()
