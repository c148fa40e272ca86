"""Golden vectors for the single-correlation visibility / weight preparation of `pfb init`
(/root/reference/src/pfb_imaging/utils/correlations.py:195-232: `_weight_data_impl`, `wgt_func`, `vis_func`), made by
EXECUTING the reference's own numba functions in this container: their definitions are cut out of the module's AST at
generation time (the module itself imports dask / xarray, which are not installed) and compiled by numba.

  weight_data_corr.npz  <- cases: complex128 / complex64 data with 2 and 4 correlations, diagonal Jones per
                           (time, antenna, channel, direction 0, correlation), ragged time bins, unit and random gains

The per-Stokes variant (`utils/weighting.py:274-468`) takes its expressions from `radiomesh.generated._stokes_expr`,
which is neither in /root/reference nor installed: it stays unpinned.

Usage:  python tests/golden/make_golden_weight_data.py
"""
import ast
import os

import numpy as np
from numba import njit

REF = "/root/reference/src/pfb_imaging/utils/correlations.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_functions():
    tree = ast.parse(open(REF).read())
    want = ("wgt_func", "vis_func", "_weight_data_impl")
    defs = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want}
    assert set(defs) == set(want)
    ns = {"np": np, "njit": njit}
    # the helpers first: the implementation calls them by name
    for name in want:
        exec(compile(ast.Module(body=[defs[name]], type_ignores=[]), REF, "exec"), ns)
    return ns["_weight_data_impl"]


def main():
    impl = reference_functions()
    rng = np.random.default_rng(20261019)
    out = {}
    k = 0
    for cdt, rdt in ((np.complex128, np.float64), (np.complex64, np.float32)):
        for ncorr in (2, 4):
            for unit in (False, True):
                nt, nant, nchan = 7, 6, 5
                counts = rng.integers(3, 12, nt)
                tbin_counts = counts.astype(np.int64)
                tbin_idx = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64) + 40  # chunk offset, as dask hands it over
                nrow = int(counts.sum())
                ant1 = rng.integers(0, nant - 1, nrow).astype(np.int32)
                ant2 = (ant1 + rng.integers(1, nant - ant1)).astype(np.int32)
                data = (rng.standard_normal((nrow, nchan, ncorr)) + 1j * rng.standard_normal((nrow, nchan, ncorr))).astype(cdt)
                weight = rng.uniform(0.2, 2.0, (nrow, nchan, ncorr)).astype(rdt)
                if unit:
                    jones = np.ones((nt, nant, nchan, 1, 2), dtype=cdt)
                else:
                    jones = (rng.uniform(0.5, 1.5, (nt, nant, nchan, 1, 2)) *
                             np.exp(1j * rng.uniform(-np.pi, np.pi, (nt, nant, nchan, 1, 2)))).astype(cdt)
                vis, wgt = impl(data, weight, jones, tbin_idx.copy(), tbin_counts, ant1, ant2)
                for name, arr in (("data", data), ("weight", weight), ("jones", jones), ("tbin_idx", tbin_idx),
                                  ("tbin_counts", tbin_counts), ("ant1", ant1), ("ant2", ant2), ("vis", vis), ("wgt", wgt)):
                    out[f"{name}_{k}"] = arr
                k += 1
    out["ncase"] = np.int64(k)
    np.savez_compressed(os.path.join(HERE, "weight_data_corr.npz"), **out)
    print("wrote", k, "cases")


if __name__ == "__main__":
    main()
