"""GPU: the fp32 pair FFT engine (csrc/fft2.cuh: packed two-wide butterflies, 128-byte XOR swizzle) against numpy, and
the plane transforms built on it — TMA-fed column passes (csrc/cols2.cuh), two-planes-per-CTA row passes
(csrc/rows2.cuh) — against the explicit DFT and against the scalar-engine kernels of fused_fft.cuh
(PFBG_COLS=old PFBG_ROWS=old) on the same inputs."""
import ctypes as C

import numpy as np
import pytest

from oracle import dft
from pfb_imaging_b200 import _lib, wgridder as W
from pfbg_testutil import rel_l2, small_problem

pytestmark = pytest.mark.gpu

# every radix (2, 3, 4, 5, 7, 8, 11, 16), every stage addressing mode (linear, M*NP == 8, M == 1, generic)
SIZES = [32, 64, 96, 160, 224, 256, 352, 480, 512, 1024, 1120, 2048, 3072, 3360, 4096, 5760, 6144, 7168]


def _first_radix(n):
    """First stage of csrc/pfbgrid.cu::factorize: odd radices first, then 16s, then the remaining power of two."""
    for r in (11, 7, 5, 3):
        if n % r == 0:
            return r
    a = (n & -n).bit_length() - 1
    return 16 if a >= 4 else 1 << a


@pytest.mark.parametrize("np_", [1, 2, 4])
@pytest.mark.parametrize("aos", [0, 1, 2])  # 0: pair-element input, DIF; 1: dense AoS input, DIF; 2: DIT
def test_pair_engine_matches_numpy(gpu, np_, aos):
    lib = _lib.load()
    rng = np.random.default_rng(11)
    for n in SIZES:
        if n * np_ * 16 > 232448:
            continue
        if aos == 1 and (n // _first_radix(n)) * np_ % 8:
            continue  # dense first-stage input needs a first stride of a multiple of 8 chunks (p2_dense_ok)
        x = (rng.standard_normal((2, 2 * np_, n)) + 1j * rng.standard_normal((2, 2 * np_, n))).astype(np.complex64)
        for inverse in (0, 1):
            out = np.empty_like(x)
            _lib.check(lib.pfbg_debug_fft2(0, n, np_, 2, C.c_void_p(x.ctypes.data), C.c_void_p(out.ctypes.data),
                                           inverse, aos))
            x64 = x.astype(np.complex128)
            ref = np.fft.ifft(x64, axis=2) * n if inverse else np.fft.fft(x64, axis=2)
            assert rel_l2(out, ref) <= 3e-6, (n, np_, aos, inverse, rel_l2(out, ref))


PFB = dict(flip_u=False, flip_v=True, flip_w=False, do_wgridding=True, divide_by_n=False)


@pytest.mark.parametrize("nx,ny,fov", [(128, 96, 0.25), (256, 256, 0.4), (192, 320, 0.1)])
def test_tma_column_transforms_against_dft_and_old_kernels(gpu, monkeypatch, nx, ny, fov):
    """nx / 2 a multiple of 32: the geometry the TMA-fed column kernels serve.  Degridding, gridding and the fused
    Hessian against the DFT, then the same calls through PFBG_COLS=old."""
    eps = 1e-5
    W.clear_plan_pool()
    p = small_problem(nrow=900, nchan=3, nx=nx, ny=ny, seed=5, wscale=2.0, fov=fov)
    kw = dict(center_x=0.01, center_y=-0.02, **PFB)
    img = p["img"].astype(np.float32)
    vis, wgt = p["vis"].astype(np.complex64), p["wgt"].astype(np.float32)
    act = p["mask"] != 0
    res = {}
    for mode in ("new", "old"):
        if mode == "old":
            monkeypatch.setenv("PFBG_COLS", "old")
            monkeypatch.setenv("PFBG_ROWS", "old")
        with W.plan_for(p["uvw"], p["freq"], npix_x=nx, npix_y=ny, pixsize_x=p["cell"], pixsize_y=p["cell"],
                        epsilon=eps, precision="single", mask=p["mask"], **kw) as gp:
            v = gp.degrid(img)
            d = gp.grid(vis, wgt)
            gp.bind_weights(wgt)
            h = gp.hessian(img, wsum=2.0, eta=0.25)
            res[mode] = (v, d, h)
        W.clear_plan_pool()
    v, d, h = res["new"]
    ref = dft.dft_dirty2vis(p["uvw"], p["freq"], p["img"], p["cell"], p["cell"], **kw)
    assert rel_l2(v[act], ref[act]) <= eps
    dref = dft.dft_vis2dirty(p["uvw"], p["freq"], vis, wgt, p["mask"], nx, ny, p["cell"], p["cell"], **kw)
    assert rel_l2(d, dref) <= eps
    href = dft.dft_vis2dirty(p["uvw"], p["freq"], ref, wgt, p["mask"], nx, ny, p["cell"], p["cell"], **kw) / 2.0 \
        + 0.25 * p["img"]
    assert rel_l2(h, href) <= 2 * eps
    vo, do, ho = res["old"]
    assert rel_l2(v[act], vo[act]) <= 3e-6 and rel_l2(d, do) <= 3e-6 and rel_l2(h, ho) <= 3e-6


def test_tma_column_transforms_partial_uv_window(gpu, monkeypatch):
    """Samples confined to a corner of the uv plane: the active window is a small circular range that wraps around
    row 0 / column 0, so the inverse pass loads two row segments and the forward pass writes a wrapped range."""
    eps = 1e-5
    nx = ny = 256
    W.clear_plan_pool()
    p = small_problem(nrow=600, nchan=2, nx=nx, ny=ny, seed=9, wscale=1.0, fov=0.2)
    uvw = p["uvw"].copy()
    uvw[:, :2] *= 0.2  # |u|, |v| < 20 % of the grid: window wraps around the origin
    uvw[:, 0] += 0.05 * np.abs(uvw[:, 0]).max()
    kw = dict(**PFB)
    img = p["img"].astype(np.float32)
    vis, wgt = p["vis"].astype(np.complex64), p["wgt"].astype(np.float32)
    out = {}
    for mode in ("new", "old"):
        if mode == "old":
            monkeypatch.setenv("PFBG_COLS", "old")
            monkeypatch.setenv("PFBG_ROWS", "old")
        with W.plan_for(uvw, p["freq"], npix_x=nx, npix_y=ny, pixsize_x=p["cell"], pixsize_y=p["cell"], epsilon=eps,
                        precision="single", **kw) as gp:
            out[mode] = (gp.degrid(img), gp.grid(vis, wgt))
        W.clear_plan_pool()
    ref = dft.dft_dirty2vis(uvw, p["freq"], p["img"], p["cell"], p["cell"], **kw)
    assert rel_l2(out["new"][0], ref) <= eps
    dref = dft.dft_vis2dirty(uvw, p["freq"], vis, wgt, None, nx, ny, p["cell"], p["cell"], **kw)
    assert rel_l2(out["new"][1], dref) <= eps
    assert rel_l2(out["new"][0], out["old"][0]) <= 3e-6 and rel_l2(out["new"][1], out["old"][1]) <= 3e-6
