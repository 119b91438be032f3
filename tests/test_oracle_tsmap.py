"""CPU checks of the TS-map oracle: the restated Brent solver against the installed scipy, the restated compiled
statistics against the reference's own Cython module (oracle/_ref), and the golden fixture."""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import cash, tsmap

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ts_map_small.npz")


def test_brentq_restatement_equals_scipy():
    scipy_optimize = pytest.importorskip("scipy.optimize")
    rng = np.random.RandomState(0)
    for case in range(200):
        a, b, c = rng.uniform(0.5, 3), rng.uniform(-2, 2), rng.uniform(0.1, 5)

        def f(x):
            return np.tanh(a * (x - b)) * c + 0.1 * (x - b) ** 3

        lo, hi = b - rng.uniform(0.1, 10), b + rng.uniform(0.1, 10)
        rtol = 10 ** rng.uniform(-12, -1)
        root, info = scipy_optimize.brentq(f, lo, hi, rtol=rtol, maxiter=100, full_output=True)
        mine = tsmap.brentq(f, lo, hi, rtol=rtol, maxiter=100)
        assert mine == (root, info.iterations, info.function_calls), case
    # the failures the reference catches (ts.py:855)
    with pytest.raises(ValueError):
        scipy_optimize.brentq(lambda x: x * x + 1, -1, 1)
    with pytest.raises(tsmap.BrentError):
        tsmap.brentq(lambda x: x * x + 1, -1, 1)
    with pytest.raises(RuntimeError):
        scipy_optimize.brentq(lambda x: np.arctan(x - 0.3), -1e6, 1e3, rtol=1e-14, maxiter=3)
    with pytest.raises(tsmap.BrentError):
        tsmap.brentq(lambda x: np.arctan(x - 0.3), -1e6, 1e3, rtol=1e-14, maxiter=3)
    assert tsmap.brentq(lambda x: x, 0.0, 1.0) == (0.0, 0, 2)     # root at a bound: no iteration


def test_restated_statistics_equal_reference_cython():
    ref = cash.reference_module()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference once)")
    rng = np.random.RandomState(3)
    for n in (1, 7, 200):
        counts = rng.poisson(2.0, size=n).astype(float)
        background = rng.uniform(0.0, 3.0, size=n)
        model = rng.uniform(0.0, 2.0, size=n) * (rng.uniform(size=n) > 0.2)
        background[rng.uniform(size=n) < 0.1] = 0.0
        for x in (-0.2, 0.0, 0.7, 13.0):
            assert tsmap.f_cash_root(x, counts, background, model) == ref.f_cash_root_cython(x, counts, background, model)
        assert_allclose(tsmap.norm_bounds(counts, background, model), ref.norm_bounds_cython(counts, background, model),
                        rtol=0, atol=0)
        npred = background + 0.7 * model
        assert_allclose(tsmap.cash_sum(counts, npred), ref.cash_sum_cython(counts, npred), rtol=1e-15)
    # no positive model bin: the Cython division by zero (cdivision) gives inf / nan bounds, not an exception
    z = np.zeros(4)
    got, want = tsmap.norm_bounds(np.ones(4), np.ones(4), z), ref.norm_bounds_cython(np.ones(4), np.ones(4), z)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and got[2] == want[2] == -1e14


def test_golden_fixture_reproduces():
    """The committed fixture equals what the oracle computes now (reference Cython + scipy when available, else the
    restatements, which then must agree to rounding)."""
    g = np.load(GOLDEN)
    inputs = {k: g[k] for k in ("counts", "background", "exposure", "kernel", "mask")}
    sel = np.zeros_like(inputs["mask"])
    sel[8:11, 10:15] = True       # around the bright source
    sel[16:20, 18:24] = True      # the zero-counts corner
    sel[6:9, 1:4] = True          # next to the empty strip
    sel &= inputs["mask"]
    out = tsmap.ts_map(inputs["counts"], inputs["background"], inputs["exposure"], inputs["kernel"], mask=sel,
                       rtol=0.01, max_niter=100, n_sigma=1.0)
    exact = cash.reference_module() is not None
    for name in tsmap.RESULT_NAMES:
        want = np.where(sel, g[f"out_{name}"], np.nan)
        if exact:
            assert np.array_equal(out[name], want, equal_nan=True), name
        else:
            assert_allclose(out[name], want, rtol=1e-9, atol=1e-9, equal_nan=True, err_msg=name)
    # the branches the fixture is meant to cover
    assert np.nanmax(g["out_ts"]) > 100 and g["out_niter"][20, 23] == 0 and np.isnan(g["out_ts"][4, 21])
    assert np.nanmin(g["out_norm"]) < 0


def test_restatements_alone_agree_with_reference_run():
    g = np.load(GOLDEN)
    inputs = {k: g[k] for k in ("counts", "background", "exposure", "kernel")}
    for pos in ((9, 12), (15, 5), (18, 21), (20, 23), (7, 2), (0, 25)):
        res = tsmap.ts_value(pos, rtol=0.01, max_niter=100, n_sigma=1.0, use_reference=False, use_scipy=False, **inputs)
        for name in tsmap.RESULT_NAMES:
            assert_allclose(res[name], g[f"out_{name}"][pos], rtol=1e-10, atol=1e-10, err_msg=f"{name} at {pos}")
