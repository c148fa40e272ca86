"""GPU: l2 re-weighting kernel vs the reference's own statements (tests/golden/l2_reweight.npz) and the
model-transfer / residual / noise branches of image_data_products (operators/gridder.py:477-757)."""
import os

import numpy as np
import pytest

from oracle import dft
from oracle import weighting as ow
from pfb_imaging_b200 import operators as ops
from pfb_imaging_b200 import weighting as gw
from pfb_imaging_b200 import wgridder as W
from pfbg_testutil import GOLDEN, rel_l2, small_problem

pytestmark = pytest.mark.gpu


def test_l2_reweight_matches_reference(gpu):
    g = np.load(os.path.join(GOLDEN, "l2_reweight.npz"))
    for k in range(int(g["ncase"])):
        wp = g[f"wgtp_{k}"]
        wp = None if wp.size == 0 else wp
        w = g[f"wgt_{k}"].copy()
        ret = gw.l2_reweight(g[f"rv_{k}"], w, g[f"mask_{k}"], float(g[f"dof_{k}"]), wgtp=wp)
        assert ret is w  # in place, like `wgt *= ...` (gridder.py:530)
        # fp64: the sum over samples is the only reordered operation; fp32 data: ressq is float32 in both
        tol = 1e-12 if w.dtype == np.float64 else 2e-6
        np.testing.assert_allclose(w, g[f"out_{k}"], rtol=tol, atol=0)


def test_l2_reweight_edges(gpu):
    # exactly zero residuals: None, weights untouched (gridder.py:531-532)
    w = np.ones((1, 8, 3))
    assert gw.l2_reweight(np.zeros((1, 8, 3), np.complex128), w, np.ones((8, 3), np.uint8), 2.0) is None
    assert np.array_equal(w, np.ones((1, 8, 3)))
    # two correlations, ragged size (not a multiple of the block), no mask array
    rng = np.random.default_rng(3)
    rv = rng.standard_normal((2, 1237, 3)) + 1j * rng.standard_normal((2, 1237, 3))
    rv[1] *= 3.0
    w = rng.uniform(0.5, 1.5, rv.shape)
    want = ow.l2_reweight(rv, w, np.ones((1237, 3), np.uint8), 4.0)
    got = gw.l2_reweight(rv, w, None, 4.0)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    # empty selection
    e = np.zeros((1, 0, 3))
    assert gw.l2_reweight(e.astype(np.complex128), e, np.zeros((0, 3), np.uint8), 2.0) is not None
    # dtype coupling is enforced
    with pytest.raises(ValueError):
        gw.l2_reweight(rv.astype(np.complex64), w, None, 4.0)


@pytest.mark.parametrize("robust", [None, 0.0])
def test_image_data_products_model_l2_residual(gpu, robust):
    p = small_problem(nrow=600, nchan=3, nx=64, ny=80, seed=11)
    nx, ny, cell = 64, 80, p["cell"]
    l0, m0 = 0.01, -0.02
    fu, fv, fw, x0, y0 = ops.wgridder_conventions(l0, m0)
    model = np.zeros((1, nx, ny))
    model[0, 20, 30], model[0, 40, 11], model[0, 5, 70] = 1.0, 0.5, 2.0
    vis, wgt, mask = p["vis"][None], p["wgt"][None], p["mask"]
    wgt0 = wgt.copy()
    dof = 3.0
    out = ops.image_data_products_arrays(p["uvw"], p["freq"], vis, wgt, mask, nx, ny, 96, 120, cell, cell, l0=l0, m0=m0,
                                         epsilon=1e-8, model=model, l2_reweight_dof=dof, robustness=robust,
                                         do_noise=True, rng=7)
    assert np.array_equal(wgt, wgt0)  # the caller's weights are left alone
    # model visibilities against the explicit DFT (the reference's own ground truth), unmasked as at :485-503
    rows = np.arange(0, 600, 7)
    mv = dft.dft_dirty2vis(p["uvw"], p["freq"], model[0], cell, cell, x0, y0, fu, fv, fw, True, False, rows=rows)
    rv_rows = vis[0][rows] - mv
    # expected weights from the oracle chain (l2 block, then the imaging weights)
    com = dict(uvw=p["uvw"], freq=p["freq"], pixsize_x=cell, pixsize_y=cell, center_x=x0, center_y=y0, epsilon=1e-8,
               flip_u=fu, flip_v=fv, flip_w=fw, divide_by_n=False, sigma_min=1.1, sigma_max=3.0)
    rv = vis[0] - W.dirty2vis(dirty=model[0], **com)
    assert rel_l2(rv[rows], rv_rows) <= 1e-7
    want_w = ow.l2_reweight(rv[None], wgt, mask, dof)
    if robust is not None:
        nxp, nyp = int(np.ceil(1.7 * nx)), int(np.ceil(1.7 * ny))
        nxp, nyp = nxp + nxp % 2, nyp + nyp % 2
        cnt = ow.compute_counts(p["uvw"], p["freq"], mask, want_w, nxp, nyp, cell, cell, np.float64, 1, -1.0, 1.0)
        cnt = ow.box_sum_counts(ow.filter_extreme_counts(cnt, level=5.0), 0)
        want_w = ow.counts_to_weights(cnt, p["uvw"], p["freq"], want_w, mask, nxp, nyp, cell, cell, robust, -1.0, 1.0)
    sel = mask.astype(bool)
    np.testing.assert_allclose(out["weight"][:, sel], want_w[:, sel], rtol=1e-6, atol=0)
    np.testing.assert_allclose(out["wsum"], want_w[:, sel].sum(axis=-1), rtol=1e-7)
    # images: the same gridder calls with those weights; and residual = dirty - R^H W R model
    ww = np.ascontiguousarray(out["weight"][0])
    dirty = W.vis2dirty(vis=vis[0], wgt=ww, mask=mask, npix_x=nx, npix_y=ny, **com)
    resid = W.vis2dirty(vis=rv, wgt=ww, mask=mask, npix_x=nx, npix_y=ny, **com)
    assert rel_l2(out["dirty"][0], dirty) <= 1e-12 and rel_l2(out["residual"][0], resid) <= 1e-12
    conv = W.vis2dirty(vis=W.dirty2vis(dirty=model[0], **com), wgt=ww, mask=mask, npix_x=nx, npix_y=ny, **com)
    assert rel_l2(out["residual"][0], out["dirty"][0] - conv) <= 1e-9
    px = (np.arange(0, nx, 9), np.arange(0, ny, 11))
    dref = dft.dft_vis2dirty(p["uvw"], p["freq"], rv, ww, mask, nx, ny, cell, cell, x0, y0, fu, fv, fw, True, False,
                             pixels=px)
    assert rel_l2(out["residual"][0][px], dref) <= 1e-7
    # noise: seeded, finite, and of the scale sqrt(sum w) a W^-1-covariance draw grids to
    out2 = ops.image_data_products_arrays(p["uvw"], p["freq"], vis, wgt, mask, nx, ny, 96, 120, cell, cell, l0=l0,
                                          m0=m0, epsilon=1e-8, model=model, l2_reweight_dof=dof, robustness=robust,
                                          do_noise=True, rng=7, do_psf=False, do_dirty=False)
    assert rel_l2(out2["noise"], out["noise"]) <= 1e-12 and "dirty" not in out2 and "psf" not in out2
    assert 0.3 < out["noise"].std() / np.sqrt(out["wsum"][0]) < 3.0


def test_image_data_products_l2_needs_model(gpu):
    p = small_problem(nrow=50, nchan=2, nx=32, ny=32, seed=2)
    with pytest.raises(ValueError, match="no model passed in"):
        ops.image_data_products_arrays(p["uvw"], p["freq"], p["vis"][None], p["wgt"][None], p["mask"], 32, 32, 48, 48,
                                       p["cell"], p["cell"], l2_reweight_dof=2.0)
