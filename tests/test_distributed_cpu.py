"""Host-side logic of the sharded joint fit on CPU: size-balanced partition + one all-reduce per evaluation
(gloo, world_size 2), mirroring test_map_fit_ray's "same numbers as the serial fit" (datasets/tests/test_map.py:891-954)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
from numpy.testing import assert_allclose

from gammapy_b200.distributed import ShardedDatasets, partition


def test_partition_balanced_and_complete():
    sizes = [510, 510, 510, 510, 7, 7, 7, 7, 100]
    shards = partition(sizes, 4)
    assert sorted(i for s in shards for i in s) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= 100
    assert partition([1] * 64, 8) == [list(range(r, 64, 8)) for r in range(8)]
    assert partition([5, 3], 4)[2:] == [[], []]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dataset_stat(i, values):
    """Deterministic stand-in for one dataset's Cash sum as a function of the shared parameter vector."""
    w = np.sin(np.arange(values.shape[1]) + i)
    return (values * w).sum(axis=1) ** 2 + i


def _worker(rank, world, port, n_datasets, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = partition([1 + (i % 3) for i in range(n_datasets)], world)[rank]

        def local_stat(values):
            total = np.zeros(values.shape[0])
            for i in mine:
                total += _dataset_stat(i, values)
            return total

        sharded = ShardedDatasets(local_datasets=None, local_stat=local_stat)
        values = np.linspace(0.5, 1.5, 5 * 3).reshape(5, 3)
        got = sharded.stat_sum_batch(values)
        single = sharded.stat_sum_batch(values[2])
        out.put((rank, got.tolist(), single))
    finally:
        dist.destroy_process_group()


def test_sharded_stat_sum_gloo_world2():
    world, n_datasets = 2, 7
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_datasets, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    values = np.linspace(0.5, 1.5, 5 * 3).reshape(5, 3)
    want = sum(_dataset_stat(i, values) for i in range(n_datasets))
    for rank, got, single in results:
        assert_allclose(got, want, rtol=1e-13)  # every rank holds the joint statistic
        assert_allclose(single, want[2], rtol=1e-13)


# ---------------------------------------------------------------------------------------------------------
# joint FIT over two ranks: global parameter order with per-dataset background models that live on one rank
# ---------------------------------------------------------------------------------------------------------
def _toy_models(n_datasets):
    from gammapy_b200.modeling.models import FoVBackgroundModel, PowerLawSpectralModel, PointSpatialModel, SkyModel

    shared = SkyModel(spatial_model=PointSpatialModel(lon_0=0.0, lat_0=0.0, frame="galactic"),
                      spectral_model=PowerLawSpectralModel(index=2.0, amplitude=1e-12, reference=1), name="src")
    shared.spatial_model.lon_0.frozen = True
    shared.spatial_model.lat_0.frozen = True
    backgrounds = [FoVBackgroundModel(dataset_name=f"obs-{i}") for i in range(n_datasets)]
    return shared, backgrounds


def _toy_stat(i, index, amplitude, norm):
    """Smooth stand-in for dataset i's Cash sum: minimum at index 2.3, amplitude 2e-12, norm 1 + 0.05 i."""
    return (1 + 0.1 * i) * (50 * (index - 2.3) ** 2 + 30 * (np.log(amplitude / 2e-12)) ** 2
                            + 400 * (norm - 1 - 0.05 * i) ** 2 + 10 * (index - 2.3) * (norm - 1 - 0.05 * i)) + i


def _fit_worker(rank, world, port, n_datasets, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gammapy_b200.modeling import Fit

        shared, backgrounds = _toy_models(n_datasets)
        mine = partition([1] * n_datasets, world)[rank]
        # local column order: the shared source's parameters, then the background models of the local datasets
        local_pars = list(shared.parameters)
        for i in mine:
            local_pars += list(backgrounds[i].parameters)
        names = [p.name for p in local_pars]
        i_index, i_amp = names.index("index"), names.index("amplitude")
        norm_cols = [k for k, p in enumerate(local_pars) if p.name == "norm"]

        def local_stat(values):
            total = np.zeros(values.shape[0])
            for k, i in enumerate(mine):
                total += _toy_stat(i, values[:, i_index], values[:, i_amp], values[:, norm_cols[k]])
            return total

        sharded = ShardedDatasets(None, local_stat=local_stat, global_models=[shared] + backgrounds,
                                  local_parameters=local_pars)
        result = Fit().run(sharded)
        pars = sharded.parameters
        out.put((rank, result.success, result.total_stat, [p.value for p in pars], [p.error for p in pars],
                 [p.name for p in pars]))
    finally:
        dist.destroy_process_group()


def test_sharded_joint_fit_gloo_world2():
    world, n_datasets = 2, 5
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fit_worker, args=(r, world, port, n_datasets, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, ok0, stat0, val0, err0, names), (_, ok1, stat1, val1, err1, _) = results
    assert ok0 and ok1
    assert stat0 == stat1 and val0 == val1 and err0 == err1      # both ranks made identical decisions
    got = dict(zip(names, val0))
    assert abs(got["index"] - 2.3) < 1e-3
    assert abs(got["amplitude"] / 2e-12 - 1) < 1e-2
    norms = [v for n, v in zip(names, val0) if n == "norm"]
    assert len(norms) == n_datasets
    assert_allclose(norms, [1 + 0.05 * i for i in range(n_datasets)], atol=2e-3)
    assert abs(stat0 - sum(range(n_datasets))) < 1e-3


def test_fit_confidence_on_a_quadratic_likelihood():
    """Fit.confidence (modeling/fit.py:302-348, scipy backend semantics) on the toy likelihood, single process:
    profile and fixed-parameter intervals against the analytic values of the quadratic form."""
    from gammapy_b200.modeling import Fit

    shared, backgrounds = _toy_models(1)
    local_pars = list(shared.parameters) + list(backgrounds[0].parameters)
    names = [p.name for p in local_pars]
    i_index, i_amp, i_norm = names.index("index"), names.index("amplitude"), names.index("norm")

    def local_stat(values):
        return _toy_stat(0, values[:, i_index], values[:, i_amp], values[:, i_norm])

    sharded = ShardedDatasets(None, local_stat=local_stat, global_models=[shared] + backgrounds,
                              local_parameters=local_pars)
    fit = Fit()
    result = fit.run(sharded)
    assert result.success
    index = shared.spectral_model.index
    best = index.value
    assert abs(best - 2.3) < 1e-3
    prof = fit.confidence(sharded, index, sigma=1, reoptimize=True)
    assert prof["success"] and prof["success_errn"] and prof["success_errp"]
    # 50 d^2 + 400 n^2 + 10 d n minimised over n: (50 - 1/16) d^2 = 1
    assert_allclose([prof["errn"], prof["errp"]], 1 / np.sqrt(50 - 1 / 16), rtol=2e-3)
    fixed = fit.confidence(sharded, index, sigma=2, reoptimize=False)
    assert_allclose([fixed["errn"], fixed["errp"]], 2 / np.sqrt(50), rtol=2e-3)
    assert index.value == best and not index.frozen              # status restored
    assert fixed["nfev_errp"] > 0 and abs(fixed["stat_null"]) < 1e-3
    # an interval that does not fit inside the parameter limits is reported, not invented
    index.max = best + 0.05
    clipped = fit.confidence(sharded, index, sigma=1, reoptimize=False)
    assert not clipped["success_errp"] and np.isnan(clipped["errp"]) and clipped["success_errn"]


def test_stat_surface_with_and_without_reoptimisation():
    from gammapy_b200.modeling import Fit

    shared, backgrounds = _toy_models(1)
    local_pars = list(shared.parameters) + list(backgrounds[0].parameters)
    names = [p.name for p in local_pars]
    i_index, i_amp, i_norm = names.index("index"), names.index("amplitude"), names.index("norm")
    sharded = ShardedDatasets(None, local_stat=lambda v: _toy_stat(0, v[:, i_index], v[:, i_amp], v[:, i_norm]),
                              global_models=[shared] + backgrounds, local_parameters=local_pars)
    fit = Fit()
    assert fit.run(sharded).success
    index, norm = shared.spectral_model.index, backgrounds[0].spectral_model.norm
    index.scan_values = [2.2, 2.3, 2.4]
    norm.scan_values = [0.95, 1.0]
    amp0 = shared.spectral_model.amplitude.value
    plain = fit.stat_surface(sharded, index, norm, reoptimize=False)
    assert plain["stat_scan"].shape == (3, 2)
    want = np.array([[_toy_stat(0, a, amp0, b) for b in (0.95, 1.0)] for a in (2.2, 2.3, 2.4)])
    assert_allclose(plain["stat_scan"], want, rtol=1e-9, atol=1e-9)
    shared.spectral_model.amplitude.value = 3e-12       # off the optimum: the re-fit must bring it back at every point
    refit = fit.stat_surface(sharded, index, norm, reoptimize=True)
    best = np.array([[_toy_stat(0, a, 2e-12, b) for b in (0.95, 1.0)] for a in (2.2, 2.3, 2.4)])
    assert_allclose(refit["stat_scan"], best, atol=2e-3)
    assert refit["fit_results"].shape == (3, 2) and all(r.success for r in refit["fit_results"].ravel())
    assert shared.spectral_model.amplitude.value == 3e-12 and not index.frozen and not norm.frozen


def test_priors_enter_the_joint_statistic_and_the_fit():
    """GaussianPrior / UniformPrior (modeling/models/prior.py) as -2 log pdf terms of Datasets.stat_sum
    (datasets/core.py:264-271), seen by Fit; values against scipy.stats."""
    from scipy import stats as sps

    from gammapy_b200.modeling import Fit, GaussianPrior, Parameter, Parameters, UniformPrior

    p = Parameter("x", 1.3)
    p.prior = GaussianPrior(mu=1.0, sigma=0.2)
    assert_allclose(p.prior(p), -2 * sps.norm(1.0, 0.2).logpdf(1.3), rtol=1e-13)
    q = Parameter("y", 0.4)
    q.prior = UniformPrior(min=0.0, max=2.0, weight=3)
    assert_allclose(q.prior(q), 3 * -2 * sps.uniform(0.0, 2.0).logpdf(0.4), rtol=1e-13)
    assert_allclose(Parameters([p, q, Parameter("z", 5.0)]).prior_stat_sum(), p.prior(p) + q.prior(q), rtol=1e-14)
    q.value = 2.5
    assert q.prior(q) == np.inf
    assert_allclose(p.prior.evaluate_many([1.0, 1.3]), [-2 * sps.norm(1.0, 0.2).logpdf(1.0), p.prior(p)], rtol=1e-13)

    # a Gaussian prior on the index pulls the toy fit between the likelihood optimum (2.3) and the prior mean
    shared, backgrounds = _toy_models(1)
    local_pars = list(shared.parameters) + list(backgrounds[0].parameters)
    names = [pp.name for pp in local_pars]
    i_index, i_amp, i_norm = names.index("index"), names.index("amplitude"), names.index("norm")
    sharded = ShardedDatasets(None, local_stat=lambda v: _toy_stat(0, v[:, i_index], v[:, i_amp], v[:, i_norm]),
                              global_models=[shared] + backgrounds, local_parameters=local_pars)
    index = shared.spectral_model.index
    index.prior = GaussianPrior(mu=2.0, sigma=0.1)
    result = Fit().run(sharded)
    assert result.success
    # minimise (50 - 1/16) (x - 2.3)^2 + (x - 2)^2 / 0.01: weights 49.9375 and 100
    want = (49.9375 * 2.3 + 100 * 2.0) / (49.9375 + 100)
    assert abs(index.value - want) < 2e-3
    assert_allclose(index.error, np.sqrt(1 / (49.9375 + 100)), rtol=2e-2)
    assert_allclose(sharded.stat_sum(), result.total_stat, rtol=0, atol=1e-6)


def test_parameter_estimator_on_the_toy_likelihood():
    """ParameterEstimator (estimators/parameter.py): best fit, delta(TS) against a null value, errn / errp, upper
    limit and scan of the background norm, against the closed forms of the quadratic toy likelihood."""
    from gammapy_b200.estimators import ParameterEstimator

    shared, backgrounds = _toy_models(1)
    local_pars = list(shared.parameters) + list(backgrounds[0].parameters)
    names = [p.name for p in local_pars]
    i_index, i_amp, i_norm = names.index("index"), names.index("amplitude"), names.index("norm")
    sharded = ShardedDatasets(None, local_stat=lambda v: _toy_stat(0, v[:, i_index], v[:, i_amp], v[:, i_norm]),
                              global_models=[shared] + backgrounds, local_parameters=local_pars)
    norm = backgrounds[0].spectral_model.norm
    norm.scan_values = [0.9, 1.0, 1.1]
    norm.min, norm.max = 0.0, 3.0
    est = ParameterEstimator(n_sigma=1, n_sigma_ul=2, null_value=0.5, selection_optional="all", reoptimize=True)
    res = est.run(sharded, norm)
    # profile in the norm: 400 n^2 + 10 d n + 50 d^2 minimised over d -> (400 - 0.5) n^2, n = norm - 1
    curv = 400 - 0.5
    assert res["success"] and abs(res["norm"] - 1.0) < 1e-3
    assert_allclose(res["norm_err"], 1 / np.sqrt(curv), rtol=2e-2)
    assert_allclose(res["ts"], curv * 0.25, rtol=2e-3)
    assert_allclose([res["norm_errn"], res["norm_errp"]], 1 / np.sqrt(curv), rtol=3e-3)
    assert_allclose(res["norm_ul"], 1.0 + 2 / np.sqrt(curv), rtol=3e-3)
    assert_allclose(res["stat_scan"] - res["stat"], curv * np.array([0.01, 0.0, 0.01]), atol=2e-3)
    assert not norm.frozen and norm.value == 1.0              # status restored (the value before the run)
    fixed = ParameterEstimator(null_value=0.5, reoptimize=False).run(sharded, norm)
    assert fixed["ts"] > 0 and "norm_ul" not in fixed
