"""GPU: the reference-level drop-ins added in round 2 — BandWorkerPool (operators/band_worker.py:209-319),
HessianTree / HessTreeRay (operators/hessian.py:439-615), compute_residual (operators/gridder.py:1019-1148),
image_data_products (:375-757), _comps2vis_impl (:276-367), pcg_dds (opt/pcg.py:444-583), the direct-mode PSF
Hessian (operators/hessian.py:178-248) and the device solvers against iterates produced by the reference's own
numba solvers (tests/golden/solvers.npz)."""
import ctypes as C
import inspect
import os

import numpy as np
import pytest

from oracle import dft
from pfb_imaging_b200 import operators as ops, psf as P, solvers, store
from pfbg_testutil import GOLDEN, rel_l2, small_problem
from test_store_cpu import make_dt_store

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, "solvers.npz"))


def test_hessian_tree_matches_the_reference_class(gpu):
    parts = [dict(psfhat=G[f"ht_psfhat{p}"], beam=G[f"ht_beam{p}"], wsum=G[f"ht_wsum{p}"]) for p in range(3)]
    x = G["ht_x"]
    ncorr, nx, ny = x.shape
    nxp, nyo2 = parts[0]["psfhat"].shape[1:]
    ht = P.HessianTree(parts, nx, ny, nxp, 2 * (nyo2 - 1), eta=0.3, nthreads=1)
    assert len(ht._groups) == 2 * ncorr  # partitions 0 and 1 share a beam: merged by linearity
    assert rel_l2(ht.dot(x), G["ht_dot"]) <= 1e-12
    assert rel_l2(ht.hdot(x), G["ht_dot"]) <= 1e-12
    ht.close()
    ht = P.HessianTree(parts, nx, ny, nxp, 2 * (nyo2 - 1), eta=0.1, nthreads=1, wsum=17.0)
    assert rel_l2(ht.dot(x), G["ht_dot_wsum"]) <= 1e-12
    ht.close()
    with pytest.raises(ValueError):
        P.HessianTree([], nx, ny, nxp, 2 * (nyo2 - 1))


def _dev_matvec(a):
    """apply_dev(in_ptr, out_ptr, stream) for a dense host matrix: device -> host, matvec, host -> device."""
    import torch

    rt = C.CDLL("libcudart.so")
    n = a.shape[0]

    def apply_dev(ip, op, stream):
        torch.cuda.synchronize()
        h = np.empty(n)
        assert rt.cudaMemcpy(C.c_void_p(h.ctypes.data), ip, C.c_size_t(8 * n), 2) == 0
        o = a @ h
        assert rt.cudaMemcpy(op, C.c_void_p(o.ctypes.data), C.c_size_t(8 * n), 1) == 0

    return apply_dev


def test_device_solvers_reproduce_reference_iterates(gpu):
    a, b, x0 = G["pcg_a"], G["pcg_b"], G["pcg_x0"]
    ap = _dev_matvec(a)
    for k in (1, 2, 5, 12):
        x = solvers.pcg_device(ap, b, x0=x0.copy(), tol=0.0, maxit=k, minit=k, verbosity=0)
        np.testing.assert_allclose(x, G[f"pcg_numba_k{k}"], rtol=1e-9, atol=1e-11)
    x0b = x0.copy()
    assert solvers.pcg_device(ap, b, x0=x0b, tol=0.0, maxit=3, minit=3, verbosity=0) is x0b  # x0 is bound as the iterate
    for k in (1, 4, 25):
        beta, bv = solvers.power_method_device(ap, b.shape, b0=G["pm_b0"].copy(), tol=0.0, maxit=k, verbosity=0)
        np.testing.assert_allclose(beta, G[f"pm_numba_beta_k{k}"], rtol=1e-10)
        np.testing.assert_allclose(bv, G[f"pm_numba_b_k{k}"], rtol=1e-8, atol=1e-11)


def test_band_worker_pool_is_a_drop_in(gpu, tmp_path):
    # same constructor / method signatures as the reference class
    want = {"__init__": ["self", "nband", "nthreads"], "load_bands": ["self", "store_url", "node_names"],
            "init_hess": ["self", "partitions_per_band", "nx", "ny", "nx_psf", "ny_psf", "etas", "wsums"],
            "hess_dot": ["self", "x"], "hess_cg": ["self", "rhs", "x0", "tol", "maxit", "minit", "verbosity"],
            "init_psi": ["self", "nx", "ny", "bases", "nlevel"], "psi_dot": ["self", "x", "alphao"],
            "psi_hdot": ["self", "alpha", "xo"],
            "residual": ["self", "model", "cell_rad", "epsilon", "do_wgridding", "double_accum"], "get_mem": ["self"]}
    for name, args in want.items():
        assert list(inspect.signature(getattr(ops.BandWorkerPool, name)).parameters) == args, name
    nband, nx, nxp = 2, 32, 48
    path = str(tmp_path / "img.dt")
    truth = make_dt_store(path, nband=nband, npart=2, nx=nx, nx_psf=nxp)
    pool = ops.BandWorkerPool(nband, nthreads=2)
    pool.load_bands(path, [f"band{b:04d}" for b in range(nband)])
    with pytest.raises(ValueError):
        pool.load_bands(path, ["band0000"])
    # Hessian role from the loaded partitions, total-wsum convention
    etas, wsums = np.array([0.2, 0.3]), np.array([50.0, 50.0])
    pool.init_hess(None, nx, nx, nxp, nxp, etas, wsums)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((nband, 1, nx, nx))
    hx = pool.hess_dot(x[:, 0])
    assert hx.shape == (nband, nx, nx)
    for b in range(nband):
        ref = np.zeros((nx, nx))
        for t in truth[b]["parts"]:
            xpad = np.zeros((nxp, nxp))
            xpad[:nx, :nx] = x[b, 0] * t["BEAM"][0]
            ref += t["BEAM"][0] * np.fft.irfft2(np.fft.rfft2(xpad) * np.abs(t["PSFHAT"][0]), s=(nxp, nxp))[:nx, :nx]
        ref = ref / wsums[b] + etas[b] * x[b, 0]
        assert rel_l2(hx[b], ref) <= 1e-12
    sol = pool.hess_cg(hx, None, 1e-10, 300, 1, 0)
    assert rel_l2(sol, x[:, 0]) <= 1e-6
    # Psi role
    nxmax, nymax = pool.init_psi(nx, nx, ["self", "db1", "db2"], 2)
    alpha = np.zeros((nband, 3, nxmax, nymax))
    pool.psi_dot(x[:, 0], alpha)
    xo = np.zeros((nband, nx, nx))
    pool.psi_hdot(alpha, xo)
    assert rel_l2(xo, 3.0 * x[:, 0]) <= 1e-12  # three orthonormal bases: Psi Psi^H = 3 I
    # exact residual role: (nband, corr, nx, ny) in and out, reference argument order
    model = np.zeros((nband, 1, nx, nx))
    model[:, 0, 10, 12] = 1.0
    model[1, 0, 20, 5] = 0.5
    cell = truth[0]["parts"][0]["cell"]
    res = pool.residual(model, cell, 1e-7, True, True)
    assert res.shape == (nband, 1, nx, nx)
    for b in range(nband):
        conv = np.zeros((nx, nx))
        for t in truth[b]["parts"]:
            mv = dft.dft_dirty2vis(t["UVW"], t["FREQ"], t["BEAM"][0] * model[b, 0], cell, cell, flip_v=True, divide_by_n=False)
            conv += dft.dft_vis2dirty(t["UVW"], t["FREQ"], mv, t["WEIGHT"][0], t["MASK"], nx, nx, cell, cell, flip_v=True,
                                      divide_by_n=False)
        assert rel_l2(truth[b]["dirty"][0] - res[b, 0], conv) <= 2e-7
    assert np.array_equal(pool.residual(np.zeros_like(model), cell), np.stack([truth[b]["dirty"] for b in range(nband)]))
    assert len(pool.get_mem()) == nband
    # the facade the deconvolution drivers use
    ht = P.HessTreeRay(None, nx, nx, nxp, nxp, etas=etas, wsums=wsums, workers=pool)
    assert rel_l2(ht.dot(x[:, 0]), hx) <= 1e-13 and rel_l2(ht.cg(hx, tol=1e-10, maxit=300), x[:, 0]) <= 1e-6
    pool.close()
    ops.clear_plan_cache()


def _dds(tmp_path, seed=0, nx=48, unit_beam=False):
    p = small_problem(nrow=800, nchan=3, nx=nx, ny=nx, seed=seed)
    rng = np.random.default_rng(seed + 1)
    beam = np.ones((1, nx, nx)) if unit_beam else rng.uniform(0.7, 1.0, (1, nx, nx))
    wsum = float(p["wgt"][p["mask"] > 0].sum())
    truth = np.zeros((nx, nx))
    truth[nx // 2, nx // 2], truth[10, 30] = 2.0, 1.0
    kw = dict(flip_v=True, divide_by_n=False)
    mv = dft.dft_dirty2vis(p["uvw"], p["freq"], beam[0] * truth, p["cell"], p["cell"], **kw)
    dirty = dft.dft_vis2dirty(p["uvw"], p["freq"], mv, p["wgt"], p["mask"], nx, nx, p["cell"], p["cell"], **kw)
    ds = store.Dataset(attrs=dict(flip_u=False, flip_v=True, flip_w=False, x0=0.0, y0=0.0, wsum=wsum, cell_rad=p["cell"],
                                  bandid=3))
    ds["UVW"], ds["FREQ"] = (("row", "three"), p["uvw"]), (("chan",), p["freq"])
    ds["MASK"], ds["BEAM"] = (("row", "chan"), p["mask"]), (("corr", "x", "y"), beam)
    path = str(tmp_path / "band.dds")
    return p, beam, dirty, truth, wsum, ds, path


def test_compute_residual_reads_and_writes_the_store(gpu, tmp_path):
    p, beam, dirty, truth, wsum, ds, path = _dds(tmp_path)
    nx = truth.shape[0]
    ds["WEIGHT"] = (("corr", "row", "chan"), p["wgt"][None])
    ds["DIRTY"] = (("corr", "x", "y"), dirty[None])
    ds.to_zarr(path, mode="w")
    assert list(inspect.signature(ops.compute_residual).parameters) == [
        "dsl", "nx", "ny", "cellx", "celly", "output_name", "model", "nthreads", "epsilon", "do_wgridding", "double_accum",
        "verbosity", "async_write"]
    model = truth[None] * 0.5
    res, fut = ops.compute_residual([path], nx, nx, p["cell"], p["cell"], path, model, nthreads=1, epsilon=1e-8)
    assert fut is not None and fut.done() and res.shape == (1, nx, nx)
    assert rel_l2(res[0], 0.5 * dirty) <= 1e-7  # the operator is linear: half the true model leaves half the dirty image
    back = store.open_zarr(path)
    assert np.array_equal(back.RESIDUAL.values, res) and np.array_equal(back.MODEL.values, model)
    assert np.array_equal(back.DIRTY.values, dirty[None])  # untouched
    res2, fut2 = ops.compute_residual(path, nx, nx, p["cell"], p["cell"], path, model, async_write=False)
    assert fut2 is None and rel_l2(res2, res) <= 1e-6
    ops.clear_plan_cache()


def test_pcg_dds_mops_the_flux_with_the_exact_hessian(gpu, tmp_path):
    # (unit beam: the reference's final residual applies the beam to the model term twice, opt/pcg.py:556-570 through
    # hessian.py:94-95, so only then is it the residual of the system that was solved)
    p, beam, dirty, truth, wsum, ds, path = _dds(tmp_path, seed=4, unit_beam=True)
    ds["WEIGHT"] = (("row", "chan"), p["wgt"])  # pcg_dds works on one correlation (opt/pcg.py:521)
    ds["BEAM"] = (("x", "y"), beam[0])
    ds["DIRTY"] = (("x", "y"), dirty)
    ds.to_zarr(path, mode="w")
    assert list(inspect.signature(solvers.pcg_dds).parameters)[:6] == ["ds_name", "eta", "mask", "use_psf", "residual_name",
                                                                      "model_name"]
    mask = np.zeros_like(truth)
    mask[truth > 0] = 1.0
    resid, bandid = solvers.pcg_dds(path, 1e-6, mask=mask, epsilon=1e-8, tol=1e-9, maxit=100, verbosity=0)
    assert bandid == 3
    back = store.open_zarr(path)
    model = back.MODEL_MOPPED.values
    # sources are recovered where the mask allows flux; the residual is what the exact operator leaves
    assert np.allclose(model[truth > 0], truth[truth > 0], rtol=1e-3)
    again = ops.hessian_slice(model, uvw=p["uvw"], weight=p["wgt"], vis_mask=p["mask"], freq=p["freq"], beam=beam[0],
                              cell=p["cell"], epsilon=1e-8)
    assert np.abs((dirty - again) - resid).max() <= 1e-6 * np.abs(dirty).max()
    assert np.abs(resid).max() <= 1e-3 * np.abs(dirty).max()
    assert np.array_equal(back.RESIDUAL_MOPPED.values, resid) and "UPDATE" in back and "X0" in back
    ops.clear_plan_cache()


def test_comps2vis_impl_against_the_dft(gpu):
    p = small_problem(nrow=300, nchan=6, nx=48, ny=48, seed=6)
    rng = np.random.default_rng(2)
    ncomp = 5
    locx, locy = rng.integers(0, 48, ncomp), rng.integers(0, 48, ncomp)
    coeffs = rng.uniform(0.5, 2.0, (2, ncomp))  # flux = c0 + c1 * f
    mds = store.Dataset(attrs=dict(cell_rad_x=p["cell"], cell_rad_y=p["cell"], npix_x=48, npix_y=48, center_x=0.0,
                                   center_y=0.0, flip_u=False, flip_v=True, flip_w=False))
    mds["coefficients"] = (("par", "comps"), coeffs)
    mds["location_x"], mds["location_y"] = (("comps",), locx), (("comps",), locy)
    modelf = lambda t, f, c0, c1: c0 + c1 * f  # noqa: E731
    tfunc, ffunc = (lambda t: t), (lambda f: f / 1e9 - 1.0)
    utime = np.array([10.0])
    one = np.array([0])
    fbi, fbc = np.array([0, 2]), np.array([2, 4])  # two imaging bands: channels [0:2], [2:6]
    vis = ops._comps2vis_impl(p["uvw"], utime, p["freq"], one, np.array([300]), one, np.array([1]), fbi, fbc,
                              np.ones((48, 48), bool), mds, modelf, tfunc, ffunc, epsilon=1e-8, product="I")
    assert vis.shape == (300, 6, 1) and vis.dtype == np.complex128
    for sl in (slice(0, 2), slice(2, 6)):
        img = np.zeros((48, 48))
        img[locx, locy] = modelf(10.0, ffunc(p["freq"][sl].mean()), *coeffs)
        ref = dft.dft_dirty2vis(p["uvw"], p["freq"][sl], img, p["cell"], p["cell"], flip_v=True, divide_by_n=False)
        assert rel_l2(vis[:, sl, 0], ref) <= 1e-8
    # channels outside the requested range stay zero
    v2 = ops._comps2vis_impl(p["uvw"], utime, p["freq"], one, np.array([300]), one, np.array([1]), fbi, fbc,
                             np.ones((48, 48), bool), mds, modelf, tfunc, ffunc, epsilon=1e-8, freq_max=p["freq"][1] + 1.0)
    assert not v2[:, 2:].any() and rel_l2(v2[:, :2], vis[:, :2]) <= 1e-12


def test_image_data_products_writes_the_band_group(gpu, tmp_path):
    p = small_problem(nrow=500, nchan=3, nx=40, ny=40, seed=8)
    parts = []
    for sl in (slice(0, 200), slice(200, 500)):
        d = store.Dataset()
        d["UVW"], d["FREQ"] = (("row", "three"), p["uvw"][sl]), (("chan",), p["freq"])
        d["VIS"] = (("corr", "row", "chan"), p["vis"][None, sl])
        d["WEIGHT"], d["MASK"] = (("corr", "row", "chan"), p["wgt"][None, sl]), (("row", "chan"), p["mask"][sl])
        d["BEAM"] = (("corr", "l_beam", "m_beam"), np.ones((1, 4, 4)))
        d["l_beam"], d["m_beam"] = (("l_beam",), np.linspace(-1, 1, 4)), (("m_beam",), np.linspace(-1, 1, 4))
        parts.append(d)
    out_path = str(tmp_path / "band.dds")
    assert list(inspect.signature(ops.image_data_products).parameters)[:10] == [
        "dsl", "dsp", "nx", "ny", "nx_psf", "ny_psf", "cellx", "celly", "output_name", "attrs"]
    outs = ops.image_data_products(parts, None, 40, 40, 56, 56, p["cell"], p["cell"], out_path, dict(timeid=7, bandid=1),
                                   robustness=0.0, do_beam=True, epsilon=1e-8)
    assert outs["timeid"] == 7 and outs["psf"].shape == (1, 56, 56)
    ref = ops.image_data_products_arrays(p["uvw"], p["freq"], p["vis"][None], p["wgt"][None], p["mask"], 40, 40, 56, 56,
                                         p["cell"], p["cell"], robustness=0.0, epsilon=1e-8)
    assert rel_l2(outs["residual"], ref["dirty"]) <= 1e-9 and rel_l2(outs["psf"], ref["psf"]) <= 1e-9
    ds = store.open_zarr(out_path)
    for k in ("DIRTY", "PSF", "PSFHAT", "WSUM", "WEIGHT", "UVW", "MASK", "FREQ", "BEAM"):
        assert k in ds, k
    assert np.allclose(ds.WSUM.values, ref["wsum"]) and ds.flip_v is True and ds.bandid == 1
    assert rel_l2(ds.PSFHAT.values, ref["psfhat"]) <= 1e-9 and np.allclose(ds.BEAM.values, 1.0)
    W = __import__("pfb_imaging_b200.wgridder", fromlist=["x"])
    W.clear_plan_pool()


def test_direct_mode_and_cube_convolutions(gpu):
    rng = np.random.default_rng(12)
    nband, nx, ny, nxp, nyp = 2, 48, 40, 64, 56
    psf = np.zeros((nband, nxp, nyp))
    psf[:, nxp // 2 - 3: nxp // 2 + 4, nyp // 2 - 3: nyp // 2 + 4] = np.outer(np.hanning(7), np.hanning(7))
    psfhat = np.fft.rfft2(np.fft.ifftshift(psf, axes=(1, 2)), axes=(1, 2))
    abspsf = np.abs(psfhat)
    x = rng.standard_normal((nband, nx, ny))
    xout = np.empty_like(x)
    P.psf_convolve_cube(None, None, xout, psfhat, nyp, x)
    for b in range(nband):
        ref = np.fft.irfft2(np.fft.rfft2(np.pad(x[b], ((0, nxp - nx), (0, nyp - ny)))) * psfhat[b], s=(nxp, nyp))[:nx, :ny]
        assert rel_l2(xout[b], ref) <= 1e-12
    x4 = x[:, None]
    x4o = np.empty_like(x4)
    P.psf_convolve_fscube(None, None, x4o, psfhat[:, None], nyp, x4)
    assert rel_l2(x4o[:, 0], xout) <= 1e-13
    taper = P.taperf((nx, ny), 8)
    for mode in ("forward", "backward"):
        got = P.hess_direct_slice(x[0], abspsf=abspsf[0], taperxy=taper, lastsize=nyp, eta=0.7, mode=mode)
        k = abspsf[0] + 0.7
        k = k if mode == "forward" else 1.0 / k
        ref = taper * np.fft.irfft2(np.fft.rfft2(np.pad(x[0] * taper, ((0, nxp - nx), (0, nyp - ny)))) * k, s=(nxp, nyp))[:nx, :ny]
        assert rel_l2(got, ref) <= 1e-12
    H = P.HessPSF(nx, ny, abspsf, beam=None, eta=0.5, cgtol=1e-9, cgmaxit=300, cgverbose=0, taper_width=8)
    d = H.idot(x, mode="direct")
    for b in range(nband):
        assert rel_l2(d[b], P.hess_direct_slice(x[b], abspsf=abspsf[b], taperxy=taper, lastsize=nyp,
                                                eta=0.5 * np.sqrt(nx * ny), mode="backward")) <= 1e-13
    hx = H.dot(x)
    assert rel_l2(H.idot(hx, mode="psf"), x) <= 1e-5  # CG from the direct estimate
    with pytest.raises(ValueError):
        H.idot(x, mode="nope")
    H.close()
    P.clear_convolver_cache()


def test_band_pool_on_shared_plane_stacks(gpu):
    """The bands of a GPU take turns on lent plane stacks (pfbg_plan_set_stack, wgridder.StackArena): one or two
    stacks for three bands give the results of three owned stacks, through the pipelined pool call and band by band;
    a plan without a stack refuses to transform."""
    from pfb_imaging_b200 import wgridder as Wg

    p = small_problem(nrow=600, nchan=4, nx=96, ny=64, seed=77, wscale=2.0)
    rng = np.random.default_rng(1)
    freqs = [p["freq"] * f for f in (1.0, 1.1, 1.25)]
    x = rng.standard_normal((3, 96, 64)).astype(np.float32)

    def make(external):
        return {b: ops.BandHessian(p["uvw"], fr, p["wgt"].astype(np.float32), p["mask"], 96, 64, p["cell"], epsilon=1e-5,
                                   precision="single", wsum=3.0, external_stack=external) for b, fr in enumerate(freqs)}

    own = ops.BandPool(make(False), nband=3)
    ref = own.hess_dot(x)
    own.close()
    lone = make(True)[0]
    with pytest.raises(RuntimeError, match="no plane stack"):
        lone.dot(x[0])
    lone.close()
    for nslots in (1, 2):
        pool = ops.BandPool(make(True), nband=3, share_stacks=nslots)
        assert pool._arena.nslots == nslots and len(pool._arena.buffers) == nslots
        got = pool.hess_dot(x)
        assert rel_l2(got, ref) <= 1e-6
        for b in range(3):
            assert rel_l2(pool.ops[b].dot(x[b]), ref[b]) <= 1e-6
        pool.close()
    # plans that own a stack give it up when one is lent
    pool = ops.BandPool(make(False), nband=3, share_stacks=True)
    assert rel_l2(pool.hess_dot(x), ref) <= 1e-6
    pool.close()
    with pytest.raises(RuntimeError, match="lent stack holds"):
        gp = Wg.plan_for(p["uvw"], p["freq"], npix_x=96, npix_y=64, pixsize_x=p["cell"], pixsize_y=p["cell"], epsilon=1e-5,
                         precision="single", external_stack=True)
        try:
            import torch

            small = torch.empty(4096, dtype=torch.uint8, device="cuda")
            gp.set_stack(small.data_ptr(), small.numel(), keep=small)
        finally:
            gp.close()
