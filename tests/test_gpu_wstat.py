"""SURVEY.md 8f N4 (statistic half): WStat / MapDatasetOnOff on the GPU (wstat_sum_kernel, gpy_wstat_sum_device) against
the reference's goldens and the oracle restatement (oracle/wstat.py, pinned by the same goldens on the CPU)."""
import json
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu
KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wstat_kats.json")))


def test_wstat_kernel_goldens_and_corner_cases(ctx):
    from gammapy_b200.stats import wstat_sum_cuda
    from oracle import wstat as ow

    total, arr = wstat_sum_cuda(KATS["n_on"], KATS["n_off"], KATS["alpha"], KATS["mu_sig"], return_array=True)
    assert_allclose(arr, KATS["wstat"])
    assert_allclose(total, np.sum(KATS["wstat"]))
    d = KATS["dataset"]
    assert_allclose(wstat_sum_cuda(d["counts"], d["counts_off"], d["alpha"], d["npred_signal"]), d["stat_sum_nomask"])
    assert_allclose(wstat_sum_cuda(d["counts"], d["counts_off"], d["alpha"], d["npred_signal"], mask=d["mask"]),
                    d["stat_sum_mask"])
    # corner cases of the profile-likelihood background + a random field, bin by bin against the oracle
    rng = np.random.RandomState(4)
    n = 200000
    mu = rng.gamma(2.0, 3.0, n)
    n_on = rng.poisson(mu + 2.0).astype(float)
    n_off = rng.poisson(8.0, n).astype(float)
    alpha = rng.uniform(0.05, 1.0, n)
    n_on[:1000] = 0
    n_off[500:2000] = 0
    mu[1500:2500] = 0.0      # mu_sig = 0: log(alpha mu_bkg), mu_bkg = 0 when n_off = 0 -> nan_to_num
    alpha[3000:3100] = 0.0   # no OFF acceptance: division by zero -> nan_to_num
    want = ow.stat_array_dataset(n_on, n_off, alpha, mu)
    total, arr = wstat_sum_cuda(n_on, n_off, alpha, mu, return_array=True)
    big = np.abs(want) > 1e300
    np.testing.assert_array_equal(np.abs(arr[big]) > 1e300, True)
    assert_allclose(arr[~big], want[~big], rtol=1e-12, atol=1e-11)
    mask = rng.uniform(size=n) < 0.7
    mask[big] = False
    assert_allclose(wstat_sum_cuda(n_on, n_off, alpha, mu, mask=mask), want[mask].sum(), rtol=1e-12)


def test_map_dataset_onoff_vs_oracle_and_fit(ctx):
    from gammapy_b200 import Fit, Map, MapDatasetOnOff, configs
    from oracle import adapter, wstat as ow

    base = configs.make_c1(width=(2.0, 2.0), mask_radius=0.9)
    src = [m for m in base.models if m.name == "model-simu"]
    de = adapter.evaluator_for(base, conv="exact64")
    sig = de.npred_signal()
    rng = np.random.RandomState(8)
    bkg = np.asarray(base.background.data, float)
    alpha = 0.2
    n_on = rng.poisson(sig + bkg).astype(float)
    n_off = rng.poisson(bkg / alpha).astype(float)
    geom = base._geom
    ds = MapDatasetOnOff(counts=Map.from_geom(geom, data=n_on, dtype=float),
                         counts_off=Map.from_geom(geom, data=n_off, dtype=float),
                         acceptance=Map.from_geom(geom, data=np.full(geom.data_shape, 1.0), dtype=float),
                         acceptance_off=Map.from_geom(geom, data=np.full(geom.data_shape, 1.0 / alpha), dtype=float),
                         exposure=base.exposure, psf=base.psf, edisp=base.edisp, mask_safe=base.mask_safe,
                         mask_fit=base.mask_fit, models=src, name="onoff")
    mask = np.asarray(base.mask.data, bool)
    want = ow.stat_sum_dataset(n_on, n_off, np.full(geom.data_shape, alpha), sig, mask)
    got = ds.stat_sum()
    assert abs(got - want) < 1e-3 + 1e-11 * abs(want), (got, want)
    arr = ds.stat_array()
    assert_allclose(arr, ow.stat_array_dataset(n_on, n_off, alpha, sig), rtol=1e-5, atol=1e-6)
    assert_allclose(ds.npred_background().data, alpha * np.nan_to_num(ow.get_wstat_mu_bkg(n_on, n_off, alpha, sig)),
                    rtol=1e-5, atol=1e-9)
    # Spectral fit with WStat as the statistic (Fit drives MapDatasetOnOff through its own engine).  The minimum is NOT
    # at the simulated parameters: most high-energy bins have n_off = 0, where the profile-likelihood background is 0
    # and WStat is biased (a property of the statistic -- the surface equals the oracle's everywhere, see above); so the
    # test checks that the optimiser lands on a minimum of that surface.
    m = src[0]
    for p in (m.spatial_model.lon_0, m.spatial_model.lat_0, m.spatial_model.sigma):
        p.frozen = True
    at_truth = ds.stat_sum()
    m.spectral_model.amplitude.value = 1.4e-11
    m.spectral_model.index.value = 2.8
    res = Fit().run(ds)
    assert res.success, res.optimize_result.message
    amp, idx = m.spectral_model.amplitude, m.spectral_model.index
    assert res.total_stat < at_truth
    assert abs(res.total_stat - ds.stat_sum()) < 1e-6
    best = (amp.value, idx.value)
    for da, di in [(1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1)]:
        amp.value, idx.value = best[0] + 0.5 * da * amp.error, best[1] + 0.5 * di * idx.error
        assert ds.stat_sum() > res.total_stat - 1e-4
    # ... and the oracle's statistic at the fitted parameters is the same number
    amp.value, idx.value = best
    base.models["model-simu"].spectral_model.amplitude.value = best[0]
    base.models["model-simu"].spectral_model.index.value = best[1]
    adapter.refresh_models(de, base)
    want = ow.stat_sum_dataset(n_on, n_off, np.full(geom.data_shape, alpha), de.npred_signal(), mask)
    assert abs(res.total_stat - want) < 1e-3


def test_joint_fit_of_a_cash_and_an_on_off_dataset(ctx):
    """``Datasets([MapDataset, MapDatasetOnOff])`` with a shared source (datasets/core.py:264-277: the sum of every
    dataset's own statistic): the Cash member through the fold kernel's statistic, the ON/OFF member through WStat;
    single and batched evaluations and a joint fit."""
    from gammapy_b200 import Datasets, Fit, Map, MapDatasetOnOff, configs
    from oracle import adapter

    base = configs.make_c1(width=(2.0, 2.0), mask_radius=0.9)
    src = [m for m in base.models if m.name == "model-simu"][0]
    de = adapter.evaluator_for(base, conv="exact64")
    base.counts = Map.from_geom(base._geom, data=de.fake(3), dtype=float)
    sig = de.npred_signal()
    rng = np.random.RandomState(8)
    bkg = np.asarray(base.background.data, float)
    alpha = 0.2
    geom = base._geom
    onoff = MapDatasetOnOff(counts=Map.from_geom(geom, data=rng.poisson(sig + bkg).astype(float), dtype=float),
                            counts_off=Map.from_geom(geom, data=rng.poisson(bkg / alpha).astype(float), dtype=float),
                            acceptance=Map.from_geom(geom, data=np.full(geom.data_shape, 1.0), dtype=float),
                            acceptance_off=Map.from_geom(geom, data=np.full(geom.data_shape, 1.0 / alpha), dtype=float),
                            exposure=base.exposure, psf=base.psf, edisp=base.edisp, mask_safe=base.mask_safe,
                            mask_fit=base.mask_fit, models=[src], name="onoff")
    joint = Datasets([base, onoff])
    assert len(joint.parameters) == len(base.parameters)  # the ON/OFF member adds no parameter: it shares the source
    want = base.stat_sum() + onoff.stat_sum()
    assert joint.stat_sum() == pytest.approx(want, rel=0, abs=1e-6)
    total, per = joint._get_engine().stat_sum(per_dataset=True)
    assert per[0] == pytest.approx(base.stat_sum(), abs=1e-6) and per[1] == pytest.approx(onoff.stat_sum(), abs=1e-6)
    vals, pars = joint.parameter_vector()
    col = [i for i, p in enumerate(pars) if p is src.spectral_model.index][0]
    grid = np.tile(vals, (3, 1))
    grid[:, col] += [0.0, 0.1, -0.1]
    scan = joint.stat_sum_batch(grid)
    assert scan[0] == pytest.approx(want, abs=1e-6) and scan[1] != scan[0] and scan[2] != scan[0]
    src.spectral_model.index.value += 0.1
    assert joint.stat_sum() == pytest.approx(scan[1], abs=1e-6)
    src.spectral_model.index.value -= 0.1
    for p in (src.spatial_model.lon_0, src.spatial_model.lat_0, src.spatial_model.sigma):
        p.frozen = True
    src.spectral_model.amplitude.value *= 1.3
    res = Fit().run(joint)
    assert res.success, res.optimize_result.message
    assert res.total_stat < want and abs(res.total_stat - joint.stat_sum()) < 1e-6
    assert abs(res.total_stat - (base.stat_sum() + onoff.stat_sum())) < 1e-6
