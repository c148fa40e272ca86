"""GPU: the in-shared-memory mixed-radix FFT engine (csrc/fft.cuh) against numpy, and the fused
plane transforms against the cuFFT path of the same library."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from pfb_imaging_b200 import _lib
from pfbg_testutil import rel_l2

pytestmark = pytest.mark.gpu

SIZES = [32, 64, 96, 160, 224, 256, 480, 512, 1120, 2048, 3360, 4096, 6144, 7168]


@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("mode", [0, 1])
def test_fft_engine_matches_numpy(gpu, prec, mode):
    lib = _lib.load()
    rng = np.random.default_rng(7)
    cdt = np.complex64 if prec == "single" else np.complex128
    tol = 3e-6 if prec == "single" else 1e-13
    for n in SIZES:
        if prec == "double" and n * 16 > 232448:
            continue
        x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(cdt)
        for inverse in (0, 1):
            out = np.empty_like(x)
            _lib.check(lib.pfbg_debug_fft1d(_lib.PFBG_F32 if prec == "single" else _lib.PFBG_F64, 0, n, 3,
                                            C.c_void_p(x.ctypes.data), C.c_void_p(out.ctypes.data), mode, inverse))
            ref = np.fft.ifft(x.astype(np.complex128), axis=1) * n if inverse else np.fft.fft(x.astype(np.complex128), axis=1)
            assert rel_l2(out, ref) <= tol, (n, mode, inverse, rel_l2(out, ref))


CHILD = """
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from pfb_imaging_b200 import wgridder as W
from pfbg_testutil import small_problem
p = small_problem(nrow=800, nchan=3, nx=96, ny=64, seed=4)
out = {{}}
for prec, eps in (("single", 1e-4), ("double", 1e-8)):
    rdt, cdt = (np.float32, np.complex64) if prec == "single" else (np.float64, np.complex128)
    gp = W.plan_for(p["uvw"], p["freq"], npix_x=96, npix_y=64, pixsize_x=p["cell"], pixsize_y=p["cell"], epsilon=eps,
                    precision=prec, mask=p["mask"], center_x=0.01, flip_v=True)
    out[prec + "_v"] = gp.degrid(p["img"].astype(rdt))
    out[prec + "_d"] = gp.grid(p["vis"].astype(cdt), p["wgt"].astype(rdt))
    gp.bind_weights(p["wgt"].astype(rdt))
    out[prec + "_h"] = gp.hessian(p["img"].astype(rdt), beam=np.full((96, 64), 0.9, rdt), wsum=3.0, eta=0.1)
    gp.close()
np.savez(sys.argv[1], **out)
"""


def test_fused_transforms_match_cufft_path(gpu, tmp_path):
    """Same inputs through PFBG_FFT=cufft (k_img2grid/k_grid2img + cuFFT) and the fused kernels."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "child.py"
    script.write_text(CHILD.format(root=root))
    res = {}
    for mode in ("cufft", "fused"):
        env = dict(os.environ, PFBG_FFT=mode)
        f = tmp_path / f"{mode}.npz"
        r = subprocess.run([sys.executable, str(script), str(f)], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        res[mode] = np.load(f)
    for k in res["cufft"].files:
        tol = 2e-5 if k.startswith("single") else 1e-12
        assert rel_l2(res["fused"][k], res["cufft"][k]) <= tol, k
