"""TS-map kernel and the two remaining compiled-stat registry entries against the oracle (reference Cython +
scipy brentq) on identical inputs.  The GPU sums each cutout by lanes + butterfly instead of serially, so the
Brent iterates agree to rounding: same iteration count, root equal far inside rtol."""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ts_map_small.npz")


def _compare(got, want, fitted, rtol_root):
    same_iter = got["niter"][fitted] == want["niter"][fitted]
    assert same_iter.mean() > 0.99, f"iteration counts differ on {100 * (1 - same_iter.mean()):.2f} % of the pixels"
    assert np.array_equal(got["success"][fitted], want["success"][fitted] != 0)
    for name in ("norm", "norm_err", "npred", "npred_excess"):
        a, b = got[name][fitted], want[name][fitted]
        both_inf = np.isinf(a) & np.isinf(b)
        assert_allclose(a[same_iter & ~both_inf], b[same_iter & ~both_inf], rtol=1e-9, atol=1e-12, err_msg=name)
        # where the solver took another number of steps both roots still satisfy the solver's tolerance
        assert_allclose(a[~same_iter & ~both_inf], b[~same_iter & ~both_inf], rtol=10 * rtol_root, atol=1e-9,
                        err_msg=name)
    for name in ("stat", "stat_null"):
        assert_allclose(got[name][fitted], want[name][fitted], rtol=1e-12, atol=1e-9, err_msg=name)
    assert_allclose(got["ts"][fitted][same_iter], want["ts"][fitted][same_iter], rtol=1e-9, atol=1e-8)


@pytest.mark.parametrize("key, rtol", [("out", 0.01), ("tight", 1e-10)])
def test_ts_map_golden(ctx, key, rtol):
    from gammapy_b200.estimators import RESULT_NAMES, compute_ts_map

    g = np.load(GOLDEN)
    got = compute_ts_map(g["counts"], g["background"], g["exposure"], g["kernel"], mask=g["mask"], rtol=rtol,
                         max_niter=100, n_sigma=1.0)
    want = {name: g[f"{key}_{name}"] for name in RESULT_NAMES}
    fitted = g["mask"].astype(bool)
    for name in RESULT_NAMES:
        if name != "success":
            assert np.isnan(got[name][~fitted]).all()
    assert not got["success"][~fitted].any()
    _compare(got, want, fitted, rtol)
    # the branches of estimate_best_fit: zero-counts window (no solver call), negative best-fit norms
    assert got["niter"][20, 23] == 0 and got["success"][20, 23]
    assert np.nanmin(got["norm"]) < 0 and np.nanmax(got["ts"]) > 100


def test_ts_map_random_field_vs_live_oracle(ctx):
    """Another field (one plane, 9 x 5 kernel, no mask, sources at the border) against the oracle run here."""
    from gammapy_b200.estimators import RESULT_NAMES, TSMapEstimator
    from oracle import tsmap

    rng = np.random.RandomState(11)
    ny, nx = 17, 19
    ky, kx = 9, 5
    gy = np.exp(-0.5 * ((np.arange(ky) - ky // 2) / 1.5) ** 2)
    gx = np.exp(-0.5 * ((np.arange(kx) - kx // 2) / 0.9) ** 2)
    kernel = (np.outer(gy, gx) / np.outer(gy, gx).sum())[None]
    exposure = rng.uniform(0.5, 1.5, size=(1, ny, nx)) * 500
    background = rng.uniform(0.2, 4.0, size=(1, ny, nx))
    counts = rng.poisson(background).astype(float)
    counts[0, 0, 0] += 40
    counts[0, 8, 9] += 25
    maps = {"counts": counts, "background": background, "exposure": exposure, "kernel": kernel,
            "mask": np.ones((1, ny, nx), dtype=bool)}
    got = TSMapEstimator(rtol=1e-3, max_niter=50, n_sigma=2).estimate_flux_map(maps)
    want = tsmap.ts_map(counts, background, exposure, kernel, mask=None, rtol=1e-3, max_niter=50, n_sigma=2.0)
    fitted = np.ones((ny, nx), dtype=bool)
    _compare(got, want, fitted, 1e-3)
    assert_allclose(got["sqrt_ts"][0, 0], np.sqrt(want["ts"][0, 0]), rtol=1e-9)
    # max_niter too small: the reference catches the solver's RuntimeError -> norm_min_total, niter = max, failure
    few = TSMapEstimator(rtol=1e-12, max_niter=2).estimate_flux_map(maps)
    ref = tsmap.ts_map(counts, background, exposure, kernel, mask=None, rtol=1e-12, max_niter=2, n_sigma=1.0)
    assert np.array_equal(few["success"], ref["success"] != 0) and not few["success"].all()
    assert_allclose(few["norm"], ref["norm"], rtol=1e-12)
    assert np.array_equal(few["niter"], ref["niter"])


def test_registry_is_complete_and_matches_reference(ctx):
    """utils/compilation.py:35-87: the cuda backend serves every key of the cython / jit dicts."""
    from gammapy_b200.stats import get_fit_statistics_compiled
    from oracle import cash, tsmap

    stats = get_fit_statistics_compiled("cuda")
    assert set(stats) == {"TRUNCATION_VALUE", "cash_sum_compiled", "weighted_cash_sum_compiled", "f_cash_root_compiled",
                          "norm_bounds_compiled"}
    ref = cash.reference_module()
    f_ref = ref.f_cash_root_cython if ref is not None else tsmap.f_cash_root
    b_ref = ref.norm_bounds_cython if ref is not None else tsmap.norm_bounds
    rng = np.random.RandomState(5)
    for n in (1, 50, 5000):
        counts = rng.poisson(1.5, size=n).astype(float)
        background = rng.uniform(0.0, 3.0, size=n)
        model = rng.uniform(0.0, 2.0, size=n) * (rng.uniform(size=n) > 0.2)
        for x in (-0.1, 0.0, 0.9, 20.0):
            want = f_ref(x, counts, background, model)
            assert_allclose(stats["f_cash_root_compiled"](x, counts, background, model), want, rtol=1e-12,
                            atol=1e-12 * model.sum())
        assert_allclose(stats["norm_bounds_compiled"](counts, background, model), b_ref(counts, background, model),
                        rtol=1e-12)
    # ties: the first bin attaining the smallest background / model ratio supplies c_min
    counts = np.array([3.0, 5.0, 7.0, 2.0])
    background = np.array([1.0, 0.5, 0.5, 4.0])
    model = np.array([1.0, 1.0, 1.0, 1.0])
    assert_allclose(stats["norm_bounds_compiled"](counts, background, model), b_ref(counts, background, model), rtol=1e-15)
    assert stats["norm_bounds_compiled"](counts, background, model)[0] == 5.0 / 4.0 - 0.5


def test_shared_memory_and_global_memory_kernels_agree(ctx, monkeypatch):
    """Small kernels take the strip kernel (cutouts staged in shared memory), big ones the any-shape kernel;
    GPY_TS_GLOBAL forces the latter: both must give the same bits (same lane dealing, same reduction order)."""
    from gammapy_b200.estimators import RESULT_NAMES, compute_ts_map

    g = np.load(GOLDEN)
    args = (g["counts"], g["background"], g["exposure"], g["kernel"])
    staged = compute_ts_map(*args, mask=g["mask"], rtol=1e-6)
    monkeypatch.setenv("GPY_TS_GLOBAL", "1")
    plain = compute_ts_map(*args, mask=g["mask"], rtol=1e-6)
    for name in RESULT_NAMES:
        assert np.array_equal(staged[name], plain[name], equal_nan=True), name
