"""Parity of the CUDA path against the oracle AT THE HEADLINE SHAPE (30 reco / 60 true energy bins):

* the full C3 dataset of BASELINE.json configs[2] (2000 x 500 px, 50 sources, 3e7 bins) -- the numpy oracle takes a
  few seconds for it, so the fused kernel is compared bin by bin with both oracle variants and the statistic with
  the reference's compiled Cash on the oracle's own npred;
* a small 30 / 60-bin dataset built to reach the corners of the fold kernel (every true-energy k-block, all four reco
  chunks, more than four sources on one tile, a diagonal-response edisp group, a ragged mask, a masked-out source);
* two observations of the c4-large geometry (1000 x 250 px, 30 / 60 bins) as a joint Datasets sum.

Tolerances as everywhere (BASELINE.json north_star): per-bin npred 1e-5 relative, statistic 1e-3 absolute.
Reference: datasets/evaluator.py:267-397, datasets/utils.py:66-84, stats/fit_statistics_cython.pyx:55-85.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu

EPS32 = np.finfo(np.float32).eps


def _longdouble_cash(counts, npred, mask=None):
    """The plain Cash formula summed in extended precision (independent of every compiled module)."""
    mu = np.where(npred > 1e-25, npred, 1e-25)
    log_mu = np.where(npred > 1e-25, np.log(mu), np.log(np.float64(np.float32(1e-25))))
    terms = mu - np.where(counts > 0, counts * log_mu, 0.0)
    if mask is not None:
        terms = terms[np.asarray(mask, bool)]
    return 2.0 * float(np.sum(terms.astype(np.longdouble)))


def _check_bins(got, de_fft, de_exact):
    want_exact = de_exact.npred()
    want_fft = de_fft.npred()
    bkg = de_exact.npred_background()
    sig = want_exact - (bkg if bkg is not None else 0.0)
    assert got.shape == want_exact.shape
    # exact-arithmetic oracle: pure relative tolerance
    assert_allclose(got, want_exact, rtol=1e-5, atol=1e-12 * float(want_exact.max()))
    # the reference's float32 FFT: 1e-5 relative plus the FFT's own noise floor per energy slice
    floor = 4 * EPS32 * sig.reshape(sig.shape[0], -1).max(axis=1)[:, None, None]
    assert np.all(np.abs(got - want_fft) <= 1e-5 * np.abs(want_fft) + floor)
    return want_exact


def test_c3_full_size_against_the_oracle(ctx):
    """BASELINE.json configs[2] exactly as bench.py times it."""
    from gammapy_b200 import Map, configs
    from oracle import adapter

    ds = configs.make_c3()
    assert ds.data_shape == (30, 500, 2000) and ds.exposure.data.shape[0] == 60
    de_fft = adapter.evaluator_for(ds, conv="fft32")
    de_exact = adapter.evaluator_for(ds, conv="exact64")
    got = ds.npred().data
    want_exact = _check_bins(got, de_fft, de_exact)
    rel = np.max(np.abs(got - want_exact) / np.maximum(want_exact, 1e-300))
    # every source's cutout equals the oracle's (a wrong slice would already fail above; this names the culprit)
    for name, ev in de_exact.evaluators.items():
        st = ds.evaluator_state(name)
        ys, xs = ev.slices
        assert (st["x0"], st["y0"], st["cx"], st["cy"]) == (xs.start, ys.start, xs.stop - xs.start, ys.stop - ys.start), name

    counts = de_fft.fake(2)     # the reference's RandomState draw from the reference-arithmetic npred
    de_exact.ds.counts = counts
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    got_stat = ds.stat_sum_likelihood()
    ref_fft, ref_exact = de_fft.stat_sum(), de_exact.stat_sum()
    exact_ld = _longdouble_cash(counts, want_exact)
    print(f"C3 full size: npred max rel err {rel:.2e}; stat gpu {got_stat!r} oracle(exact64, cython) {ref_exact!r} "
          f"oracle(fft32, cython) {ref_fft!r} longdouble {exact_ld!r}")
    # Against the exact-arithmetic oracle npred summed in extended precision: the plain 1e-3 budget.
    assert abs(got_stat - exact_ld) < 1e-3, (got_stat, exact_ld)
    # The reference sums 3e7 terms serially in one double (fit_statistics_cython.pyx:70-85); at |stat| ~ 1e8 that
    # accumulation alone is ~2e-3 from the exact sum, so the budget against it is its own rounding + 1e-3.
    assert abs(got_stat - ref_exact) < abs(ref_exact - exact_ld) + 1e-3, (got_stat, ref_exact, exact_ld)
    # The float32-FFT oracle (the reference's own arithmetic) is itself |ref_fft - exact| away from exact arithmetic
    # (FFT rounding noise in every convolved cube); the GPU must be no further from it than that plus the budget.
    assert abs(got_stat - ref_fft) < abs(ref_fft - exact_ld) + 1e-3, (got_stat, ref_fft, exact_ld)


def _patch_diagonal_edisp(de, ds, names):
    """The oracle evaluator applies ds.edisp to every model; sources that opt out get the diagonal response
    (evaluator.py:245-250) through a patched update."""
    from gammapy_b200.irf import EDispKernel

    e_true = ds.exposure.geom.axes["energy_true"]
    diag = EDispKernel.from_diagonal_response(e_true, ds._geom.axes["energy"]).pdf_matrix
    for name in names:
        ev = de.evaluators[name]
        orig = ev.update

        def update(d, ev=ev, orig=orig):
            orig(d)
            ev.edisp = diag

        ev.update = update


def _small_headline_dataset(seed=11):
    """300 x 120 px, 30 / 60 bins, 14 sources packed into 3 x 1.2 deg: many tiles see 5-9 cutouts."""
    from gammapy_b200 import Map, configs

    ds = configs.make_c3(name="c3-corners", width=(6.0, 2.4), n_sources=14)
    assert ds.data_shape == (30, 120, 300)
    rng = np.random.RandomState(seed)
    src = [m for m in ds.models if m is not ds.background_model]
    for m in src:  # pack them (make_c3 spreads over the width)
        m.spatial_model.lon_0.value = float(rng.uniform(-1.5, 1.5))
        m.spatial_model.lat_0.value = float(rng.uniform(-0.5, 0.5))
    src[3].apply_irf["edisp"] = False     # second edisp group (diagonal response)
    src[7].apply_irf["edisp"] = False
    src[5].apply_irf["psf"] = False       # image broadcast over energy
    src[9].spectral_model.amplitude.value = 0.0   # drops out (evaluator.py:385-389)
    ds.background_model.spectral_model.tilt.value = 0.13
    ds.background_model.spectral_model.norm.value = 0.93
    ds.models = ds.models
    # ragged mask: energy-dependent wedge + every 7th row off + one reco bin off entirely + safe mask hole
    nR, ny, nx = ds.data_shape
    yy, xx = np.mgrid[:ny, :nx]
    fit = np.ones(ds.data_shape, dtype=bool)
    for r in range(nR):
        fit[r] &= xx > 3 * r
        fit[r] &= (yy + r) % 7 != 0
    fit[17] = False
    safe = np.ones(ds.data_shape, dtype=bool)
    safe[:4, 40:80, 100:200] = False
    ds.mask_fit = Map.from_geom(ds._geom, data=fit, dtype=bool)
    ds.mask_safe = Map.from_geom(ds._geom, data=safe, dtype=bool)
    return ds, [src[3].name, src[7].name]


def test_headline_bins_small_dataset_corner_cases(ctx):
    from gammapy_b200 import Datasets, Map
    from oracle import adapter

    ds, diag_names = _small_headline_dataset()
    de_fft = adapter.evaluator_for(ds, conv="fft32")
    de_exact = adapter.evaluator_for(ds, conv="exact64")
    _patch_diagonal_edisp(de_fft, ds, diag_names)
    _patch_diagonal_edisp(de_exact, ds, diag_names)
    got = ds.npred().data
    want = _check_bins(got, de_fft, de_exact)
    # the field is dense: at least one tile of 128 pixels is reached by more than four cutouts
    cover = np.zeros(ds.data_shape[1:], dtype=int)
    for ev in de_exact.evaluators.values():
        if ev.contributes and ev.slices is not None:
            cover[ev.slices] += 1
    assert cover.max() > 4
    # per-model cubes (the stack order of npred_signal does not matter to a bin of one model)
    _, parts = de_exact.npred_signal(per_model=True)
    for name in diag_names + [list(de_exact.evaluators)[0], list(de_exact.evaluators)[5]]:
        part = ds.npred_signal(model_names=[name]).data
        assert_allclose(part, parts[name], rtol=1e-5, atol=1e-7 * float(parts[name].max()), err_msg=name)
    counts = de_fft.fake(3)
    de_exact.ds.counts = counts
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    stat = ds.stat_sum_likelihood()
    assert abs(stat - de_exact.stat_sum()) < 1e-3
    assert abs(stat - de_fft.stat_sum()) < 1e-3
    assert abs(stat - _longdouble_cash(counts, want, de_exact.ds.mask)) < 1e-3
    # per-bin statistic (fit_statistics.py:29-79, no n > 0 guard) against the oracle's array form
    from oracle import cash

    arr = ds.stat_array()
    want_arr = cash.cash(counts, want)
    ok = np.isfinite(want_arr)
    assert_allclose(arr[ok], want_arr[ok], rtol=1e-5, atol=1e-9)
    # a batch of vectors at this shape == one by one == oracle
    datasets = Datasets([ds])
    values, plist = datasets.parameter_vector()
    cols = [i for i, p in enumerate(plist) if p.name in ("index", "alpha")]
    grid = np.tile(values, (5, 1))
    grid[1, cols[0]] += 0.05
    grid[2, cols[3]] -= 0.07          # a diagonal-response source
    grid[3, cols] += 0.01
    grid[4, [i for i, p in enumerate(plist) if p.name == "norm"][0]] = 1.2
    batch = datasets.stat_sum_batch(grid)
    for k in range(5):
        for p, v in zip(plist, grid[k]):
            p.value = v
        one = datasets.stat_sum()
        assert one == batch[k]
        adapter.refresh_models(de_exact, ds)
        assert abs(one - de_exact.stat_sum()) < 1e-3, k
    for p, v in zip(plist, values):
        p.value = v


def test_headline_bins_source_moves_and_caches(ctx):
    """Evaluator state at 30 / 60 bins: spectral step, small move (same cutout), large move (new cutout)."""
    from gammapy_b200 import Datasets, Map
    from oracle import adapter

    ds, diag_names = _small_headline_dataset(seed=12)
    de = adapter.evaluator_for(ds, conv="exact64")
    _patch_diagonal_edisp(de, ds, diag_names)
    counts = de.fake(5)
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    datasets = Datasets([ds])
    src = [m for m in ds.models if m is not ds.background_model]

    def check(label):
        adapter.refresh_models(de, ds)
        got, want = datasets.stat_sum(), de.stat_sum()
        assert abs(got - want) < 1e-3, (label, got, want)

    check("first")
    src[0].spectral_model.parameters[0].value *= 1.01
    check("spectral")
    src[1].spatial_model.lon_0.value += 0.013
    check("small move")
    src[2].spatial_model.lat_0.value -= 0.8
    check("large move")
    src[4].spatial_model.lon_0.value = 40.0   # leaves the map: contributes -> False
    check("gone")


def test_c4_large_geometry_two_observations(ctx):
    """Two observations of the c4-large joint fit (1000 x 250 px, 30 / 60 bins), shared source, own background."""
    from gammapy_b200 import Map, configs
    from oracle import adapter, evaluator

    geometry = dict(width=(20.0, 5.0), n_reco=30, n_true=60, e_reco=(0.1, 100.0), e_true=(0.03, 300.0), mask_radius=None)
    datasets = configs.make_c4(n_obs=64, indices=[0, 37], **geometry)
    assert len(datasets) == 2 and datasets[0].data_shape == (30, 250, 1000)
    des = []
    for i, ds in enumerate(datasets):
        ds.background_model.spectral_model.norm.value = 1.0 + 0.05 * i
        de = adapter.evaluator_for(ds, conv="exact64")
        got = ds.npred().data
        want = de.npred()
        assert_allclose(got, want, rtol=1e-5, atol=1e-12 * float(want.max()))
        counts = de.fake(100 + i)
        ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
        des.append(de)
    got = datasets.stat_sum()
    want = evaluator.stat_sum(des)
    assert abs(got - want) < 1e-3, (got, want)
