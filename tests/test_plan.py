"""CPU: host-side plan logic (parameter choice, sizes, argument errors)."""
import numpy as np
import pytest

from pfb_imaging_b200 import kernel_table as kt
from pfb_imaging_b200.plan import good_size, make_plan, padded_size, w_range


def test_good_size():
    for n in (1, 2, 7, 11, 13, 97, 1000, 4097, 5734):
        g = good_size(n)
        assert g >= n
        m = g
        for p in (2, 3, 5, 7):
            while m % p == 0:
                m //= p
        assert m == 1
    assert good_size(5734) == 5760  # good_size(1.4*4096), BASELINE.md C2


def test_padded_size_is_tileable():
    for n, s, W in ((4096, 1.5, 8), (2048, 1.25, 10), (512, 2.0, 7), (100, 1.1, 16)):
        nu = padded_size(n, s, W)
        assert nu % 32 == 0 and nu >= s * n - 1e-9 and nu >= n + W


def test_kernel_table_monotone():
    for s in (1.2, 1.5, 2.0):
        errs = [kt.lookup(s, W)[1] for W in range(4, 15)]
        assert all(a > b for a, b in zip(errs, errs[1:]))


@pytest.mark.parametrize("prec,eps", [("single", 1e-5), ("single", 1e-4), ("double", 1e-7), ("double", 1e-10)])
def test_plan_meets_epsilon(prec, eps):
    p = make_plan(nx=4096, ny=4096, pixsize_x=5.5e-6, pixsize_y=5.5e-6, epsilon=eps, precision=prec,
                  wmin=-2e4, wmax=2e4, nvis=25_000_000, flip_v=True, divide_by_n=False, sigma_max=3.0)
    assert p.kernel_err <= eps / (3.0 if prec == "double" else np.sqrt(3))
    assert p.nu >= p.sigma * p.nx - 1e-9 and p.nu % 32 == 0
    assert p.nplanes >= p.W
    if prec == "single":
        assert p.W <= 8
    # every sample's support fits in the plane stack (|w| in [0, 2e4] after the Hermitian fold; planes below
    # zero are served by the mirror of planes 0..pmirror-1)
    assert p.pmirror > 0 and p.w0 == 0.5 * p.dw and p.pmirror <= p.nplanes
    for w in (0.0, 1.0, 2e4):
        ip0 = np.floor((w - p.w0) / p.dw - 0.5 * p.W) + 1
        assert -p.pmirror <= ip0 <= p.nplanes - p.W
    q = make_plan(nx=4096, ny=4096, pixsize_x=5.5e-6, pixsize_y=5.5e-6, epsilon=eps, precision=prec,
                  wmin=1.9e6, wmax=2e6, nvis=25_000_000, flip_v=True, divide_by_n=False, sigma_max=3.0)
    assert q.pmirror == 0  # far from w = 0 nothing is mirrored
    for w in (1.9e6, 2e6):
        ip0 = np.floor((w - q.w0) / q.dw - 0.5 * q.W) + 1
        assert 0 <= ip0 <= q.nplanes - q.W
    assert p.vsign == -1.0 and p.usign == 1.0


def test_plan_flip_rule_and_no_wgridding():
    p = make_plan(nx=64, ny=64, pixsize_x=1e-4, pixsize_y=1e-4, center_x=0.1, center_y=-0.2, epsilon=1e-6,
                  flip_u=True, flip_v=True, do_wgridding=False)
    assert p.center_x == -0.1 and p.center_y == 0.2 and p.nplanes == 1 and p.nshift == 0.0


def test_plan_errors():
    base = dict(nx=64, ny=64, pixsize_x=1e-4, pixsize_y=1e-4, epsilon=1e-6)
    for bad in (dict(nx=63), dict(ny=0), dict(pixsize_x=-1.0), dict(epsilon=0.0), dict(epsilon=1e-20),
                dict(precision="half"), dict(precision="single", epsilon=1e-9), dict(center_x=0.9, center_y=0.9)):
        with pytest.raises(ValueError):
            make_plan(**{**base, **bad})


def test_w_range():
    # the kernels fold w < 0 onto w > 0 (Hermitian symmetry), so the planes cover |w| only
    uvw = np.array([[0, 0, -3.0], [0, 0, 5.0]])
    f = np.array([1e9, 2e9])
    lo, hi = w_range(uvw, f)
    assert np.isclose(lo, 3.0 * 1e9 / 299792458.0) and np.isclose(hi, 5.0 * 2e9 / 299792458.0)
    lo2, hi2 = w_range(uvw, f, -1.0)
    assert (lo2, hi2) == (lo, hi)


def test_batch_plan_gives_every_snapshot_its_own_plane_block():
    from pfb_imaging_b200.plan import make_batch_plan

    wr = [(0.0, 10.0), (5.0, 400.0), (100.0, 100.0), (0.5, 60.0)]
    plan, w0, npl = make_batch_plan(wr, nx=512, ny=512, pixsize_x=2e-5, pixsize_y=2e-5, epsilon=1e-4, precision="single",
                                    sigma_min=2.0, sigma_max=2.6, divide_by_n=True)
    assert plan.pmirror == 0 and plan.sigma >= 2.0
    assert npl.dtype == np.int32 and w0.shape == (4,) and npl[2] == plan.W
    for (lo, hi), a, n in zip(wr, w0, npl):
        assert n >= plan.W and a <= lo + 1e-9 and a + (n - 1) * plan.dw >= hi - 1e-9
        # the lowest sample's first plane is >= 0 and the highest sample's last plane is < n
        assert np.floor((lo - a) / plan.dw - plan.W / 2) + 1 >= 0
        assert np.floor((hi - a) / plan.dw - plan.W / 2) + 1 + plan.W <= n
    assert plan.nplanes_std == int(npl.sum())


def test_max_stack_bytes_limits_the_plane_stack():
    """make_plan(max_stack_bytes=): the cheapest plan that fits the budget; the smallest stack when nothing fits."""
    kw = dict(nx=10240, ny=10240, pixsize_x=6e-6, pixsize_y=6e-6, epsilon=1e-5, precision="single", wmin=0.0, wmax=9000.0,
              nvis=62_500_000, sigma_min=1.1, sigma_max=3.0)
    free = make_plan(**kw)
    stack = free.nplanes * free.nu * free.nv * 8
    tight = make_plan(max_stack_bytes=int(0.9 * stack), **kw)
    assert tight.nplanes * tight.nu * tight.nv * 8 <= 0.9 * stack
    assert (tight.nu, tight.W) != (free.nu, free.W) or tight.nplanes < free.nplanes
    same = make_plan(max_stack_bytes=2 * stack, **kw)
    assert (same.nu, same.W, same.nplanes) == (free.nu, free.W, free.nplanes)
    tiny = make_plan(max_stack_bytes=1, **kw)  # nothing fits: the smallest stack the accuracy allows
    assert tiny.nplanes * tiny.nu * tiny.nv * 8 <= tight.nplanes * tight.nu * tight.nv * 8


def test_fft_cost_factor_orders_the_measured_sizes():
    """The size model of the planner against the measured order of the fp64 transform sizes (per cell
    6144 < 7168 < 5760 < 6048 < 6720) and its normalisation points."""
    from pfb_imaging_b200.plan import fft_cost_factor as f

    assert f(5760, "double") == pytest.approx(1.0) and f(6144, "single") == pytest.approx(1.0)
    assert f(6144, "double") < f(7168, "double") < f(5760, "double") < f(6720, "double")
    assert f(6144, "double") < 0.85 * f(5760, "double")
    assert f(16384, "single") < f(15360, "single")  # config 4: the power of two wins per cell
    # the C2 geometry in fp64 at eps 1e-7 lands on the 6144^2 grid (W = 11), in fp32 at 1e-5 it stays there (W = 8)
    kw = dict(nx=4096, ny=4096, pixsize_x=1.2e-5, pixsize_y=1.2e-5, wmin=0.0, wmax=2000.0, nvis=25_000_000, sigma_min=1.1,
              sigma_max=3.0)
    p64 = make_plan(epsilon=1e-7, precision="double", **kw)
    p32 = make_plan(epsilon=1e-5, precision="single", **kw)
    assert (p64.nu, p64.W) == (6144, 11) and (p32.nu, p32.W) == (6144, 8)
