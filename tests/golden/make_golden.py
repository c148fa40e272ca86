"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN
functions in this container (run here only; /root/reference does not exist on
the GPU box).  No reference source is copied into the repo: the function
definitions are parsed out of the reference files at generation time and
executed in a scratch namespace.

  kat_conventions.npz  <- explicit_degridder / explicit_wdegridder
                          (/root/reference/tests/test_hessian_approx.py:23-67) with
                          wgridder_conventions (src/pfb_imaging/operators/gridder.py:23-34)
                          on the seed-42 fixture of test_gridder_conventions (:70-125)
  weighting.npz        <- _compute_counts, counts_to_weights, filter_extreme_counts,
                          box_sum_counts (src/pfb_imaging/utils/weighting.py:81-254), numba-jitted
  uv2xy.npz            <- the uv -> cell identity of tests/test_weighting.py:113-137

Usage:  python tests/golden/make_golden.py
"""

import ast
import itertools
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path, names, ns):
    """exec the named top-level function definitions of `path` inside namespace `ns`."""
    src = open(path).read()
    tree = ast.parse(src)
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise RuntimeError(f"{path}: missing {missing}")
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, path, "exec"), ns)
    return ns


def seed42_fixture():
    # tests/test_hessian_approx.py:73-102 (same RNG call order)
    np.random.seed(42)
    npix, num_ants = 1024, 100
    pixsize = 0.5 * np.pi / 180 / 3600.0
    a1, a2 = np.asarray(list(itertools.combinations(range(num_ants), 2))).T
    antennas = 10e3 * np.random.normal(size=(num_ants, 3))
    antennas[:, 2] *= 0.001
    uvw = antennas[a1] - antennas[a2]
    freqs = np.linspace(700e6, 2000e6, 2)
    return npix, pixsize, uvw, freqs


def make_kat():
    ns = {"np": np}
    extract(f"{REF}/src/pfb_imaging/operators/gridder.py", ["wgridder_conventions"], ns)
    extract(f"{REF}/tests/test_hessian_approx.py", ["explicit_degridder", "explicit_wdegridder"], ns)
    npix, pixsize, uvw, freqs = seed42_fixture()
    rows = np.arange(0, uvw.shape[0], 25)
    sub = uvw[rows]
    offsets = [(0.0, 0.0), (0.1, -0.17), (0.2, 0.5), (-0.1, 0.2), (-0.15, -0.2)]
    out = dict(rows=rows, uvw=sub, freqs=freqs, npix=npix, pixsize=pixsize, offsets=np.array(offsets),
               flips=np.array(ns["wgridder_conventions"](0.0, 0.0)[:3], dtype=bool))
    for k, (l0, m0) in enumerate(offsets):
        # test_gridder_conventions: l = l0 + (i - n/2)(-dl)
        def lmn_a(xi, yi):
            l = l0 + (-npix / 2 + xi) * (-pixsize)
            m = m0 + (-npix / 2 + yi) * (-pixsize)
            return np.asarray([l, m, np.sqrt(1.0 - l * l - m * m)])

        lmn = [lmn_a(npix // 2, npix // 2), lmn_a(npix // 4, npix // 4)]
        for neg in (False, True):
            out[f"conv_{k}_{int(neg)}"] = ns["explicit_degridder"](sub, freqs, lmn, [1.0, 1.0], neg, convention="casa")

        # test_wgridder_conventions: l = -l0 + (i - n/2) dl ; m = m0 + (j - n/2) dm
        def lmn_b(xi, yi):
            l = -l0 + (-npix / 2 + xi) * pixsize
            m = m0 + (-npix / 2 + yi) * pixsize
            return np.asarray([l, m, np.sqrt(1.0 - l * l - m * m)])

        lmn = [lmn_b(npix // 2, npix // 2), lmn_b(npix // 4, npix // 4)]
        out[f"wconv_{k}"] = ns["explicit_wdegridder"](sub, freqs, lmn, [1.0, 1.0])
        out[f"wconv_center_{k}"] = np.array(ns["wgridder_conventions"](l0, m0)[3:])
    np.savez_compressed(os.path.join(HERE, "kat_conventions.npz"), **out)
    print("kat_conventions.npz", len(out), "arrays")


def make_weighting():
    import numba
    from numba import njit, prange
    from scipy.constants import c as lightspeed
    from scipy.ndimage import uniform_filter

    src = open(f"{REF}/src/pfb_imaging/utils/weighting.py").read()
    tree = ast.parse(src)
    want = ["_compute_counts", "counts_to_weights", "filter_extreme_counts", "box_sum_counts"]
    body = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in want)
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "JIT_OPTIONS")]
    def njit_nocache(*a, **k):  # the reference asks for an on-disk cache, which needs a real file
        k.pop("cache", None)
        return njit(*a, **k)

    ns = dict(np=np, numba=numba, njit=njit_nocache, prange=prange, lightspeed=lightspeed, uniform_filter=uniform_filter)
    exec(compile(ast.Module(body=body, type_ignores=[]), "weighting.py", "exec"), ns)

    rng = np.random.default_rng(20260101)
    nrow, nchan, ncorr = 700, 5, 2
    nx = ny = 64
    cell = 2.0e-5
    freq = np.linspace(1.0e9, 1.3e9, nchan)
    umax = 1 / cell / 2
    uvw = rng.normal(size=(nrow, 3)) * (0.45 * umax * lightspeed / freq.max())
    uvw[:5] *= 4.0  # a few samples off the grid
    mask = (rng.uniform(size=(nrow, nchan)) > 0.1).astype(np.uint8)
    out = dict(uvw=uvw, freq=freq, mask=mask, nx=nx, ny=ny, cell=cell)
    for tag, dt in (("f8", np.float64), ("f4", np.float32)):
        wgt = rng.uniform(0.5, 1.5, (ncorr, nrow, nchan)).astype(dt)
        wgt[:, rng.integers(0, nrow, 20), rng.integers(0, nchan, 20)] = 0.0
        out[f"wgt_{tag}"] = wgt
        for usign, vsign in ((-1.0, 1.0), (1.0, -1.0)):
            st = f"{tag}_{int(usign)}_{int(vsign)}"
            counts = ns["_compute_counts"](uvw, freq, mask, wgt, nx, ny, cell, cell, dt, 3, usign, vsign)
            out[f"counts_{st}"] = counts
            for robust in (-2.0, 0.0, 1.5):
                c2 = counts.copy()
                w2 = wgt.copy()
                wout = ns["counts_to_weights"](c2, uvw, freq, w2, mask, nx, ny, cell, cell, robust, usign, vsign)
                out[f"w_{st}_r{robust}"] = np.asarray(wout)
                out[f"c_{st}_r{robust}"] = c2
        cf = ns["filter_extreme_counts"](out[f"counts_{tag}_-1_1"].copy(), level=10.0)
        out[f"filtered_{tag}"] = cf
        out[f"boxsum_{tag}"] = ns["box_sum_counts"](out[f"counts_{tag}_-1_1"].copy(), 2)
    np.savez_compressed(os.path.join(HERE, "weighting.npz"), **out)
    print("weighting.npz", len(out), "arrays")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("the reference tree is needed to (re)generate golden vectors")
    make_kat()
    make_weighting()
