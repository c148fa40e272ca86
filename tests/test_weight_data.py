"""`weight_data_corr` (utils/correlations.py:195-232): the oracle restatement against vectors produced by the
reference's own numba functions (tests/golden/make_golden_weight_data.py), and the device kernel against both,
bit for bit."""
import os

import numpy as np
import pytest

from oracle import weighting as ow
from pfbg_testutil import GOLDEN


def _cases():
    z = np.load(os.path.join(GOLDEN, "weight_data_corr.npz"))
    for k in range(int(z["ncase"])):
        yield {n: z[f"{n}_{k}"] for n in ("data", "weight", "jones", "tbin_idx", "tbin_counts", "ant1", "ant2", "vis", "wgt")}


def test_oracle_matches_the_reference_numba_loop():
    n = 0
    for c in _cases():
        vis, wgt = ow.weight_data_corr(c["data"], c["weight"], c["jones"], c["tbin_idx"], c["tbin_counts"], c["ant1"], c["ant2"])
        assert vis.dtype == c["vis"].dtype and wgt.dtype == c["wgt"].dtype
        tol = 1e-6 if c["data"].dtype == np.complex64 else 1e-14  # numpy may fuse / reorder the complex products
        assert np.allclose(vis, c["vis"], rtol=tol, atol=0) and np.allclose(wgt, c["wgt"], rtol=tol, atol=0)
        n += 1
    assert n == 8


@pytest.mark.gpu
def test_device_kernel_is_bit_identical_to_the_reference(gpu):
    from pfb_imaging_b200 import weighting as wt

    for c in _cases():
        tb = c["tbin_idx"].copy()
        vis, wgt = wt.weight_data_corr(c["data"], c["weight"], c["jones"], tb, c["tbin_counts"], c["ant1"], c["ant2"])
        assert np.array_equal(tb, c["tbin_idx"])  # the caller's bins are left alone
        assert vis.dtype == c["vis"].dtype and wgt.dtype == c["wgt"].dtype
        assert np.array_equal(vis, c["vis"]) and np.array_equal(wgt, c["wgt"])


@pytest.mark.gpu
def test_rows_outside_every_bin_and_argument_checks(gpu):
    from pfb_imaging_b200 import weighting as wt

    c = next(_cases())
    counts = c["tbin_counts"].copy()
    counts[-1] -= 2  # the last two rows belong to no bin: zeros, like the reference's initialisation
    vis, wgt = wt.weight_data_corr(c["data"], c["weight"], c["jones"], c["tbin_idx"], counts, c["ant1"], c["ant2"])
    assert np.all(vis[-2:] == 0) and np.all(wgt[-2:] == 0)
    assert np.array_equal(vis[:-2], c["vis"][:-2])
    with pytest.raises(NotImplementedError):
        wt.weight_data_corr(c["data"], c["weight"], c["jones"][..., None], c["tbin_idx"], counts, c["ant1"], c["ant2"])
    with pytest.raises(ValueError):
        wt.weight_data_corr(c["data"], c["weight"][:, :1], c["jones"], c["tbin_idx"], counts, c["ant1"], c["ant2"])
