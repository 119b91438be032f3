"""Parity of the CUDA path (through the C ABI) against the oracle on identical inputs.

Tolerances (BASELINE.json north_star): per-bin npred within 1e-5 relative, total stat within 1e-3
absolute.  Against the reference's float32-FFT convolution the per-bin comparison carries the absolute
floor ``4 * eps_f32 * max(signal slice)`` (the FFT's own error is ~2e-7 x peak); against
the exact (float64) variant of the oracle no floor is needed.
"""
import ctypes as C

import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu

EPS32 = np.finfo(np.float32).eps


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------------------
# kernel 4: Cash sums (compiled-stat registry boundary)
# ---------------------------------------------------------------------------------------------------------
def test_cash_sum_registry_vs_reference(ctx):
    from gammapy_b200.stats import get_fit_statistics_compiled
    from oracle import cash

    reg = get_fit_statistics_compiled()
    assert reg["TRUNCATION_VALUE"] == 1e-25
    rng = np.random.RandomState(1)
    mu = rng.gamma(2.0, 2.0, 1_000_003)
    mu[::1000] = 0.0
    mu[5::1000] = -1.0
    mu[7::1000] = 1e-26
    n = rng.poisson(np.clip(mu, 0, None)).astype(float)
    w = rng.uniform(-0.2, 1.0, mu.size)
    ref = cash.reference_module()
    want = ref.cash_sum_cython(n, mu) if ref is not None else cash.cash_sum(n, mu)
    got = reg["cash_sum_compiled"](n, mu)
    assert abs(got - want) < 1e-6  # 1e6 bins, |stat| ~ 5e6: fp64 summation order only
    want_w = ref.weighted_cash_sum_cython(n, mu, w) if ref is not None else cash.weighted_cash_sum(n, mu, w)
    assert abs(reg["weighted_cash_sum_compiled"](n, mu, w) - want_w) < 1e-6


def test_cash_sum_edge_cases(ctx):
    from gammapy_b200.stats import cash_sum_cuda, weighted_cash_sum_cuda
    from oracle import cash

    assert cash_sum_cuda(np.zeros(0), np.zeros(0)) == 0.0
    n = np.array([0.0, 3.0, 0.0, 2.0, 1.0])
    mu = np.array([0.0, 0.0, -5.0, np.nan, 1e-30])
    assert_allclose(cash_sum_cuda(n, mu), cash.cash_sum(n, mu), rtol=1e-14)
    assert_allclose(weighted_cash_sum_cuda(n, mu, np.array([1.0, 0.0, 2.0, -1.0, 0.5])),
                    cash.weighted_cash_sum(n, mu, np.array([1.0, 0.0, 2.0, -1.0, 0.5])), rtol=1e-14)
    # golden per-bin values of the reference's test fixture
    from tests.test_oracle_kat import CASH_REF, MU_SIG, N_ON

    assert_allclose(cash_sum_cuda(np.array(N_ON, float), np.array(MU_SIG)), np.sum(CASH_REF), rtol=1e-9)


# ---------------------------------------------------------------------------------------------------------
# kernel 1: geometry + spatial models
# ---------------------------------------------------------------------------------------------------------
def _geom_desc(geom):
    from gammapy_b200 import _lib

    return _lib.GeomDesc(geom.nx, geom.ny, 1, 1, geom.binsz, geom.crval_lon, geom.crpix_x, geom.crpix_y, geom.crval_lat,
                         _lib.PROJECTIONS[geom.proj], 0)


def test_solid_angle(ctx):
    from gammapy_b200 import _lib
    from oracle import wcs

    geom = wcs.ImageGeom.create(skydir=(0.0, 0.0), binsz=0.02, width=(6, 3))
    out = np.empty(geom.data_shape)
    _lib.check(ctx.lib.gpy_solid_angle(ctx.handle, C.byref(_geom_desc(geom)), _ptr(out)))
    assert_allclose(out, geom.solid_angle(), rtol=1e-9)


SPATIAL_CASES = [
    ("point", dict(lon_0=0.013, lat_0=-0.271), 0, [0.013, -0.271]),
    ("point-center", dict(lon_0=0.0, lat_0=0.0), 0, [0.0, 0.0]),
    ("gauss-f1", dict(lon_0=0.2, lat_0=0.1, sigma=0.3, e=0.0, phi=0.0), 1, [0.2, 0.1, 0.3, 0.0, 0.0]),
    ("gauss-f2", dict(lon_0=-0.31, lat_0=0.17, sigma=0.05, e=0.0, phi=0.0), 1, [-0.31, 0.17, 0.05, 0.0, 0.0]),
    ("gauss-f7", dict(lon_0=0.4, lat_0=-0.4, sigma=0.01, e=0.0, phi=0.0), 1, [0.4, -0.4, 0.01, 0.0, 0.0]),
    ("gauss-tiny-f200", dict(lon_0=0.1, lat_0=0.1, sigma=1e-4, e=0.0, phi=0.0), 1, [0.1, 0.1, 1e-4, 0.0, 0.0]),
    ("gauss-elongated", dict(lon_0=0.1, lat_0=-0.2, sigma=0.2, e=0.7, phi=30.0), 1, [0.1, -0.2, 0.2, 0.7, 30.0]),
    ("gauss-edge", dict(lon_0=0.95, lat_0=0.0, sigma=0.05, e=0.0, phi=0.0), 1, [0.95, 0.0, 0.05, 0.0, 0.0]),
    ("gauss-outside", dict(lon_0=5.0, lat_0=0.0, sigma=0.05, e=0.0, phi=0.0), 1, [5.0, 0.0, 0.05, 0.0, 0.0]),
    ("disk", dict(lon_0=0.1, lat_0=0.2, r_0=0.3, e=0.0, phi=0.0, edge_width=0.01), 2, [0.1, 0.2, 0.3, 0.0, 0.0, 0.01]),
    ("disk-elongated", dict(lon_0=-0.2, lat_0=0.1, r_0=0.4, e=0.6, phi=110.0, edge_width=0.05), 2,
     [-0.2, 0.1, 0.4, 0.6, 110.0, 0.05]),
]


@pytest.mark.parametrize("name,pars,gtype,vec", SPATIAL_CASES, ids=[c[0] for c in SPATIAL_CASES])
def test_integrate_geom(ctx, name, pars, gtype, vec):
    from gammapy_b200 import _lib
    from oracle import spatial, wcs

    geom = wcs.ImageGeom.create(skydir=(0.0, 0.0), binsz=0.02, width=(2.0, 1.6))
    model = {"type": ["point", "gauss", "disk"][gtype], **pars}
    want = spatial.integrate_geom(model, geom).astype(np.float32)  # cast at the convolution input
    out = np.empty(geom.data_shape, dtype=np.float32)
    f = C.c_int32(0)
    v = np.array(vec, dtype=float)
    _lib.check(ctx.lib.gpy_integrate_geom(ctx.handle, C.byref(_geom_desc(geom)), gtype, 0, _ptr(v), _ptr(out), C.byref(f)))
    scale = max(float(np.abs(want).max()), 1e-300)
    assert_allclose(out, want, rtol=2e-6, atol=2 * EPS32 * scale)
    if name == "gauss-tiny-f200":
        assert f.value == 200  # test_large_oversampling: capped at MAX_OVERSAMPLING
        assert_allclose(out.sum(), 1, rtol=1e-3)
    if name == "gauss-f2":
        assert f.value == 2
    if name == "gauss-outside":
        assert f.value == 1


# ---------------------------------------------------------------------------------------------------------
# kernel 0: spectral integrals
# ---------------------------------------------------------------------------------------------------------
SPECTRAL_CASES = [
    ("pl", 0, [2.3, 4.0, 1.0], {"type": "pl", "index": 2.3, "amplitude": 4.0, "reference": 1.0}),
    ("pl-index1", 0, [1.0, 2.0, 1.0], {"type": "pl", "index": 1.0, "amplitude": 2.0, "reference": 1.0}),
    ("lp", 1, [4.0, 1.0, 2.3, 0.5], {"type": "lp", "amplitude": 4.0, "reference": 1.0, "alpha": 2.3, "beta": 0.5}),
    ("ecpl", 2, [1.6, 4.0, 1.0, 0.1, 1.0], {"type": "ecpl", "index": 1.6, "amplitude": 4.0, "reference": 1.0,
                                             "lambda_": 0.1, "alpha": 1.0}),
    ("ecpl-hard-cutoff", 2, [2.0, 1e-12, 1.0, 5.0, 1.0], {"type": "ecpl", "index": 2.0, "amplitude": 1e-12,
                                                        "reference": 1.0, "lambda_": 5.0, "alpha": 1.0}),
]


@pytest.mark.parametrize("name,gtype,vec,model", SPECTRAL_CASES, ids=[c[0] for c in SPECTRAL_CASES])
def test_spectral_integral(ctx, name, gtype, vec, model):
    from gammapy_b200 import _lib
    from oracle import spectral

    edges = np.geomspace(0.03, 300.0, 61)
    nodes = np.ascontiguousarray(spectral.integrate_nodes(edges[:-1], edges[1:]))
    want = spectral.integral(model, edges[:-1], edges[1:])
    out = np.empty(60)
    v = np.array(vec, dtype=float)
    _lib.check(ctx.lib.gpy_spectral_integral(ctx.handle, gtype, _ptr(v), _ptr(edges), 60, _ptr(nodes), nodes.shape[1],
                                             _ptr(out)))
    assert_allclose(out, want, rtol=1e-11, atol=1e-300)


def test_spectral_goldens_on_gpu(ctx):
    """modeling/models/tests/test_spectral.py TEST_MODELS integral_1_10TeV through the public model API."""
    from gammapy_b200 import ExpCutoffPowerLawSpectralModel, LogParabolaSpectralModel, PowerLawSpectralModel

    assert_allclose(PowerLawSpectralModel(index=2.3, amplitude=4, reference=1).integral(1.0, 10.0),
                    2.9227116204223784, rtol=1e-7)
    assert_allclose(ExpCutoffPowerLawSpectralModel(index=1.6, amplitude=4, reference=1, lambda_=0.1).integral(1.0, 10.0),
                    3.765838739678921, rtol=1e-5)
    assert_allclose(LogParabolaSpectralModel(alpha=2.3, amplitude=4, reference=1, beta=0.5).integral(1.0, 10.0),
                    2.255689748270628, rtol=1e-5)


# ---------------------------------------------------------------------------------------------------------
# kernel 2: PSF convolution
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,K", [((97, 113), 67), ((40, 31), 9), ((5, 7), 21), ((200, 227), 101)])
def test_psf_convolve(ctx, shape, K):
    from gammapy_b200 import _lib
    from oracle import evaluator

    rng = np.random.RandomState(3)
    img = rng.gamma(0.5, 1.0, shape).astype(np.float32)
    img[rng.uniform(size=shape) < 0.5] = 0
    yy, xx = np.mgrid[:K, :K] - (K - 1) / 2
    planes = []
    for s in (0.8, 2.0, K / 7.0):
        k = np.exp(-0.5 * (xx**2 + yy**2) / s**2) * (1 + 0.1 * xx / K)  # slightly asymmetric: catches a missing flip
        k[np.hypot(xx, yy) > 3.7 * s] = 0
        planes.append(k / k.sum())
    kern = np.array(planes)
    want = evaluator.convolve_exact64(img, kern)
    out = np.empty((3,) + shape, dtype=np.float32)
    k32 = np.ascontiguousarray(kern, dtype=np.float32)
    _lib.check(ctx.lib.gpy_psf_convolve(ctx.handle, _ptr(img), shape[0], shape[1], _ptr(k32), 3, K, _ptr(out)))
    assert_allclose(out, want, rtol=2e-6, atol=1e-7 * float(want.max()))
    ref32 = evaluator.convolve_fft32(img, kern)
    assert_allclose(out, ref32, rtol=1e-5, atol=4 * EPS32 * float(want.max()))


def test_convolve_vs_smooth_api(ctx):
    """maps/wcs/tests/test_ndmap.py:622-635 shape: Map.convolve with a normalised Gaussian kernel conserves the sum."""
    from gammapy_b200 import Map, PSFKernel, WcsGeom

    geom = WcsGeom.create(binsz=0.05, width=(5, 5))
    m = Map.from_geom(geom)
    m.data[50, 50] = 1.0
    kernel = PSFKernel.from_gauss(geom, sigma=0.1, max_radius=1.0)
    out = m.convolve(kernel)
    assert out.data.shape == m.data.shape
    assert_allclose(out.data.sum(), 1.0, rtol=1e-5)
    assert out.data[50, 50] == out.data.max()


# ---------------------------------------------------------------------------------------------------------
# full path: npred + Cash
# ---------------------------------------------------------------------------------------------------------
def _check_npred(got, de_fft, de_exact):
    want_exact = de_exact.npred()
    want_fft = de_fft.npred()
    bkg = de_exact.npred_background()
    sig = want_exact - (bkg if bkg is not None else 0.0)
    # exact arithmetic oracle: pure relative tolerance on total npred
    assert_allclose(got, want_exact, rtol=1e-5, atol=1e-12 * float(want_exact.max()))
    # the reference's float32 FFT: relative 1e-5 plus its own noise floor per energy slice
    floor = 4 * EPS32 * sig.reshape(sig.shape[0], -1).max(axis=1)[:, None, None]
    assert np.all(np.abs(got - want_fft) <= 1e-5 * np.abs(want_fft) + floor)


@pytest.mark.parametrize("width,mask_radius", [((2.0, 2.0), 0.9), ((4.0, 4.0), 1.8), ((1.5, 1.0), None)])
def test_c1_npred_and_stat(ctx, width, mask_radius):
    from gammapy_b200 import configs
    from oracle import adapter

    ds = configs.make_c1(width=width, mask_radius=mask_radius)
    de_fft = adapter.evaluator_for(ds, conv="fft32")
    de_exact = adapter.evaluator_for(ds, conv="exact64")
    got = ds.npred().data
    assert got.dtype == np.float64 and got.shape == ds.data_shape
    _check_npred(got, de_fft, de_exact)
    # cutout of the evaluator = parent slices of _cutout_view
    st = ds.evaluator_state("model-simu")
    ys, xs = de_fft.evaluators["model-simu"].slices
    assert (st["x0"], st["y0"], st["cx"], st["cy"]) == (xs.start, ys.start, xs.stop - xs.start, ys.stop - ys.start)
    # counts drawn with the reference's legacy RandomState from the oracle npred, shared by both sides
    counts = de_fft.fake(0)
    de_exact.ds.counts = counts
    from gammapy_b200 import Map

    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    stat = ds.stat_sum_likelihood()
    assert abs(stat - de_exact.stat_sum()) < 1e-3
    assert abs(stat - de_fft.stat_sum()) < 1e-3
    assert abs(stat - de_fft.stat_sum(use_reference=False)) < 1e-3


def test_npred_psf_after_edisp_total_gpu(ctx):
    """datasets/tests/test_map.py:1418-1445 on the GPU path: sum(npred) = 129553.858658."""
    from gammapy_b200 import (FoVBackgroundModel, MapAxis, MapDataset, PointSpatialModel, PowerLawSpectralModel,
                              PSFMap, SkyModel, WcsGeom)

    energy_axis = MapAxis.from_energy_bounds("1 TeV", "10 TeV", nbin=3)
    energy_axis_true = MapAxis.from_energy_bounds("0.8 TeV", "15 TeV", nbin=6, name="energy_true")
    geom = WcsGeom.create(width=4, binsz=0.02, axes=[energy_axis])
    dataset = MapDataset.create(geom=geom, energy_axis_true=energy_axis_true)
    dataset.background.data += 1
    dataset.exposure.data += 1e12
    dataset.mask_safe.data += True
    dataset.psf = PSFMap.from_gauss(energy_axis_true=energy_axis_true, sigma=0.2)
    model = SkyModel(spectral_model=PowerLawSpectralModel(), spatial_model=PointSpatialModel(), name="test-model")
    dataset.models = [FoVBackgroundModel(dataset_name=dataset.name), model]
    npred = dataset.npred()
    assert_allclose(npred.data.sum(), 129553.858658)


def test_c2_five_sources_and_batch_scan(ctx):
    from gammapy_b200 import Datasets, Map, configs
    from oracle import adapter

    ds = configs.make_c2()
    de_fft = adapter.evaluator_for(ds, conv="fft32")
    de_exact = adapter.evaluator_for(ds, conv="exact64")
    got = ds.npred().data
    _check_npred(got, de_fft, de_exact)
    # per-model signal (npred_signal(model_names=...)) against the oracle's per-evaluator cubes
    _, parts = de_exact.npred_signal(per_model=True)
    for name in ("src-0", "src-2", "src-3"):
        sig = ds.npred_signal(model_names=[name]).data
        assert_allclose(sig, parts[name], rtol=1e-5, atol=1e-7 * float(parts[name].max()))
    counts = de_fft.fake(1)
    de_exact.ds.counts = counts
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    datasets = Datasets([ds])
    base = datasets.stat_sum()
    assert abs(base - de_exact.stat_sum()) < 1e-3
    # 256-point profile of the index of source 0, all in one launch == one by one == oracle
    par = ds.models["src-0"].spectral_model.index
    values, plist = datasets.parameter_vector()
    col = [i for i, p in enumerate(plist) if p is par][0]
    scan = np.linspace(1.8, 2.8, 256)
    grid = np.tile(values, (256, 1))
    grid[:, col] = scan
    batch = datasets.stat_sum_batch(grid)
    assert batch.shape == (256,)
    for k in (0, 100, 255):
        par.value = scan[k]
        one = datasets.stat_sum()
        assert abs(one - batch[k]) < 1e-6
        adapter.refresh_models(de_exact, ds)
        assert abs(batch[k] - de_exact.stat_sum()) < 1e-3
    par.value = 2.3
    # batch with a moving spatial parameter: every vector gets its own re-convolved cube
    lon = ds.models["src-1"].spatial_model.lon_0
    col = [i for i, p in enumerate(plist) if p is lon][0]
    grid = np.tile(values, (3, 1))
    grid[:, col] += [0.0, 0.013, -0.02]
    batch = datasets.stat_sum_batch(grid)
    assert abs(batch[0] - base) < 1e-6
    for k in (1, 2):
        lon.value = grid[k, col]
        adapter.refresh_models(de_exact, ds)
        assert abs(batch[k] - de_exact.stat_sum()) < 1e-3
    lon.value = values[col]


def test_evaluator_cache_semantics(ctx):
    """datasets/tests/test_evaluator.py:242-279 spirit: spatial flux is re-evaluated only when spatial
    parameters change; the cutout is re-centred only after a drift > r_eval + 0.1 deg."""
    from gammapy_b200 import configs
    from oracle import adapter

    ds = configs.make_c1(width=(4.0, 4.0), mask_radius=1.8)
    de = adapter.evaluator_for(ds, conv="exact64")
    model = ds.models["model-simu"]
    ds.npred()
    de.npred()
    s0 = ds.evaluator_state("model-simu")
    assert s0["n_spatial_evals"] == 1 and s0["oversampling"] == 1
    model.spectral_model.index.value = 2.5
    model.spectral_model.amplitude.value = 2e-11
    ds.npred()
    assert ds.evaluator_state("model-simu")["n_spatial_evals"] == 1
    model.spatial_model.lon_0.value = 0.3  # small drift: same cutout, new convolution
    adapter.refresh_models(de, ds)
    got, want = ds.npred().data, de.npred()
    s1 = ds.evaluator_state("model-simu")
    assert s1["n_spatial_evals"] == 2
    assert (s1["x0"], s1["y0"], s1["cx"], s1["cy"]) == (s0["x0"], s0["y0"], s0["cx"], s0["cy"])
    assert_allclose(got, want, rtol=1e-5)
    model.spatial_model.sigma.value = 0.05  # r_eval shrinks to 0.25 deg: the 0.3 deg offset from (0, 0) now exceeds
    model.spatial_model.lon_0.value = 0.5   # r_eval + 0.1 -> update() re-centres the cutout
    adapter.refresh_models(de, ds)
    got, want = ds.npred().data, de.npred()
    s2 = ds.evaluator_state("model-simu")
    ys, xs = de.evaluators["model-simu"].slices
    assert (s2["x0"], s2["y0"], s2["cx"], s2["cy"]) == (xs.start, ys.start, xs.stop - xs.start, ys.stop - ys.start)
    assert (s2["cx"], s2["cy"]) != (s0["cx"], s0["cy"])
    assert s2["oversampling"] == 2
    assert_allclose(got, want, rtol=1e-5, atol=1e-12 * want.max())
    model.spectral_model.amplitude.value = 0.0  # evaluator.py:385-389
    assert_allclose(ds.npred_signal().data.sum(), 0)


def test_source_outside_mask_does_not_contribute(ctx):
    from gammapy_b200 import configs

    ds = configs.make_c1(width=(4.0, 4.0), mask_radius=0.5)
    ds.models["model-simu"].spatial_model.lon_0.value = -1.9
    ds.models["model-simu"].spatial_model.lat_0.value = -1.9
    ds.models["model-simu"].spatial_model.sigma.value = 0.02
    sig = ds.npred_signal().data
    assert sig.sum() == 0
    assert ds.evaluator_state("model-simu")["contributes"] == 0


def test_joint_datasets_sum(ctx):
    from gammapy_b200 import Map, configs
    from oracle import adapter

    datasets = configs.make_c4(n_obs=3, width=(2.0, 2.0), mask_radius=0.9)
    des = []
    for i, ds in enumerate(datasets):
        de = adapter.evaluator_for(ds, conv="exact64")
        counts = de.fake(100 + i)
        ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
        des.append(de)
    total, per = datasets._get_engine().stat_sum(per_dataset=True)
    want = [de.stat_sum() for de in des]
    assert_allclose(per, want, atol=1e-3, rtol=0)
    assert abs(total - sum(want)) < 1e-3
    assert abs(datasets.stat_sum() - total) < 1e-9


def test_errors_are_loud(ctx):
    from gammapy_b200 import GaussianSpatialModel, MapAxis, PowerLawSpectralModel, SkyModel, WcsGeom, configs

    with pytest.raises(NotImplementedError):
        WcsGeom.create(skydir=(0, 10), binsz=0.1, width=2, proj="AIT")
    ds = configs.make_dataset(width=(1.0, 1.0), mask_radius=None)
    bad = SkyModel(spatial_model=GaussianSpatialModel(frame="fk5"), spectral_model=PowerLawSpectralModel(), name="x")
    ds.models = [bad]
    with pytest.raises(NotImplementedError):
        ds.npred()
    ds.models = []
    ds.counts = None
    with pytest.raises(ValueError):
        ds.stat_sum_likelihood()
    assert MapAxis.from_energy_bounds(1, 10, 3).nbin == 3


def test_batch_with_many_spatial_tuples(ctx):
    """A finite-difference gradient batch holds more distinct spatial parameter tuples than the per-source
    cache keeps between calls: every vector must still see its own PSF-convolved cube."""
    from gammapy_b200 import Datasets, Map, configs
    from oracle import adapter

    ds = configs.make_c1(width=(2.0, 2.0), mask_radius=0.9)
    de = adapter.evaluator_for(ds, conv="exact64")
    ds.counts = Map.from_geom(ds._geom, data=de.fake(3), dtype=float)
    datasets = Datasets([ds])
    values, plist = datasets.parameter_vector()
    m = ds.models["model-simu"].spatial_model
    cols = {p.name: i for i, p in enumerate(plist)}
    grid = np.tile(values, (13, 1))
    grid[1:5, cols["lon_0"]] += [0.01, -0.01, 0.02, -0.02]
    grid[5:9, cols["lat_0"]] += [0.01, -0.01, 0.02, -0.02]
    grid[9:13, cols["sigma"]] += [0.01, -0.01, 0.02, -0.02]
    batch = datasets.stat_sum_batch(grid)
    again = datasets.stat_sum_batch(grid)
    assert_allclose(batch, again, rtol=0, atol=1e-9)  # deterministic, cache hit or not
    for k in range(13):
        m.lon_0.value, m.lat_0.value, m.sigma.value = grid[k, cols["lon_0"]], grid[k, cols["lat_0"]], grid[k, cols["sigma"]]
        adapter.refresh_models(de, ds)
        assert abs(batch[k] - de.stat_sum()) < 1e-3, k
        assert abs(batch[k] - datasets.stat_sum()) < 1e-6, k


def test_two_engines_share_one_dataset(ctx):
    """MapDataset-level calls (npred) and Datasets-level calls (stat_sum) alternate on one device handle."""
    from gammapy_b200 import Datasets, Map, configs
    from oracle import adapter

    a = configs.make_c1(name="a", width=(1.6, 1.6), mask_radius=0.7)
    b = configs.make_c2(name="b", width=(2.0, 2.0), mask_radius=0.9)
    for ds, seed in ((a, 1), (b, 2)):
        ds.counts = Map.from_geom(ds._geom, data=adapter.evaluator_for(ds, conv="exact64").fake(seed), dtype=float)
    joint = Datasets([a, b])
    want = a.stat_sum_likelihood() + b.stat_sum_likelihood()
    assert abs(joint.stat_sum() - want) < 1e-6
    npred_b = b.npred().data  # re-installs b's own parameter columns on the shared handle
    assert abs(joint.stat_sum() - want) < 1e-6
    assert_allclose(b.npred().data, npred_b)


# ---------------------------------------------------------------------------------------------------------
# shapes / options that take the other code paths
# ---------------------------------------------------------------------------------------------------------
def _stat_pair(ds, seed=5, conv="exact64"):
    from gammapy_b200 import Map
    from oracle import adapter

    de = adapter.evaluator_for(ds, conv=conv)
    ds.counts = Map.from_geom(ds._geom, data=de.fake(seed), dtype=float)
    return ds.stat_sum_likelihood(), de.stat_sum(), ds.npred().data, de.npred()


def test_odd_geometry_uses_generic_kernel(ctx):
    """nx = 101 is not a multiple of 4: the any-shape kernel runs (no bulk-copy staging, scalar loads)."""
    from gammapy_b200 import configs

    ds = configs.make_c2(width=(2.02, 1.5), mask_radius=0.7)
    assert ds._geom.npix == (101, 75)
    stat, want, npred, npred_want = _stat_pair(ds)
    assert_allclose(npred, npred_want, rtol=1e-5, atol=1e-12 * npred_want.max())
    assert abs(stat - want) < 1e-3


def test_forced_generic_kernel_matches_staged_kernel(ctx):
    """GPY_FOLD_KERNEL=2 is read at context creation; here the odd/even geometries exercise both kernels on the
    same physics: a 100-px and a 101-px wide map of the same field give the same npred where they overlap."""
    from gammapy_b200 import configs

    a = configs.make_c1(width=(2.0, 2.0), mask_radius=None)
    b = configs.make_c1(width=(2.02, 2.0), mask_radius=None)   # 101 px: generic kernel
    na, nb = a.npred().data, b.npred().data
    assert na.shape[-1] == 100 and nb.shape[-1] == 101
    # pixel centres differ by half a pixel between the two grids: compare the totals of the source-dominated region
    assert_allclose(na.sum() / nb[..., :100].sum(), 1.0, rtol=2e-2)


def test_no_psf_and_irf_switches(ctx):
    from gammapy_b200 import Map, configs
    from oracle import adapter

    ds = configs.make_c2(width=(2.0, 2.0), mask_radius=0.9)
    ds.models["src-1"].apply_irf["psf"] = False     # W broadcast over energy instead of the convolved cube
    ds.models["src-2"].apply_irf["edisp"] = False   # diagonal response: second edisp group
    ds.models = ds.models                            # rebuild evaluators (datasets/map.py:674-694)
    de = adapter.evaluator_for(ds, conv="exact64")
    for name, ev in de.evaluators.items():
        if name == "src-2":
            from gammapy_b200.irf import EDispKernel

            e_true = ds.exposure.geom.axes["energy_true"]
            ev_diag = EDispKernel.from_diagonal_response(e_true, ds._geom.axes["energy"]).pdf_matrix
            ev._diag = ev_diag
    # the oracle evaluator applies ds.edisp to every model: patch the one that opted out
    orig_update = type(de.evaluators["src-2"]).update

    def update(self, d):
        orig_update(self, d)
        if self.model["name"] == "src-2":
            self.edisp = self._diag

    type(de.evaluators["src-2"]).update = update
    try:
        want = de.npred()
    finally:
        type(de.evaluators["src-2"]).update = orig_update
    got = ds.npred().data
    assert_allclose(got, want, rtol=1e-5, atol=1e-12 * want.max())
    # dataset without any PSF
    ds2 = configs.make_c1(width=(1.6, 1.6), mask_radius=0.7)
    ds2.psf = None
    stat, want_stat, npred, npred_want = _stat_pair(ds2)
    assert_allclose(npred, npred_want, rtol=1e-6, atol=1e-12 * npred_want.max())
    assert abs(stat - want_stat) < 1e-3


def test_masks_and_missing_maps(ctx):
    from gammapy_b200 import Map, configs
    from oracle import adapter

    ds = configs.make_c1(width=(1.6, 1.6), mask_radius=0.7)
    ds.counts = Map.from_geom(ds._geom, data=adapter.evaluator_for(ds).fake(2), dtype=float)
    # everything masked: empty sum
    ds.mask_fit = Map.from_geom(ds._geom, data=np.zeros(ds.data_shape, dtype=bool), dtype=bool)
    assert ds.stat_sum_likelihood() == 0.0
    # ragged mask: one energy plane and a few scattered pixels
    rng = np.random.RandomState(0)
    mask = rng.uniform(size=ds.data_shape) < 0.3
    mask[3] = False
    ds.mask_fit = Map.from_geom(ds._geom, data=mask, dtype=bool)
    de = adapter.evaluator_for(ds, conv="exact64")
    assert abs(ds.stat_sum_likelihood() - de.stat_sum()) < 1e-3
    # no masks at all, no background map
    ds.mask_fit = None
    ds.mask_safe = None
    ds.background = None
    ds.models = [m for m in ds.models if m.name == "model-simu"]
    de = adapter.evaluator_for(ds, conv="exact64")
    assert_allclose(ds.npred().data, de.npred(), rtol=1e-5, atol=1e-12)
    assert abs(ds.stat_sum_likelihood() - de.stat_sum()) < 1e-3
    # negative amplitude: npred clipped at 0, truncation value in the Cash log (finite, large)
    ds.models["model-simu"].spectral_model.amplitude.value = -1e-11
    de = adapter.evaluator_for(ds, conv="exact64")
    got, want = ds.stat_sum_likelihood(), de.stat_sum()
    assert np.isfinite(got) and abs(got - want) < 1e-3 * max(1.0, abs(want) * 1e-9)
    # NaN parameter: the source vanishes (WcsNDMap.stack nan_to_num), background-less npred is all zero
    ds.models["model-simu"].spectral_model.amplitude.value = 1e-11
    ds.models["model-simu"].spectral_model.index.value = np.nan
    assert np.all(ds.npred().data == 0)


def test_batched_joint_datasets(ctx):
    """B > 1 with D > 1 in one launch equals the per-vector, per-dataset evaluations."""
    from gammapy_b200 import Map, configs
    from oracle import adapter

    datasets = configs.make_c4(n_obs=3, width=(1.6, 1.6), mask_radius=0.7)
    for i, ds in enumerate(datasets):
        ds.counts = Map.from_geom(ds._geom, data=adapter.evaluator_for(ds).fake(10 + i), dtype=float)
    values, plist = datasets.parameter_vector()
    cols = {id(p): i for i, p in enumerate(plist)}
    shared = datasets[0].models["model-simu"]
    grid = np.tile(values, (4, 1))
    grid[1, cols[id(shared.spectral_model.index)]] += 0.1
    grid[2, cols[id(shared.spatial_model.sigma)]] += 0.02
    grid[3, cols[id(datasets[1].background_model.spectral_model.norm)]] *= 1.1
    total, per = datasets.stat_sum_batch(grid, per_dataset=True)
    assert per.shape == (4, 3)
    assert_allclose(per.sum(axis=1), total, rtol=1e-14)
    for k in range(4):
        for p, v in zip(plist, grid[k]):
            p.value = v
        assert abs(datasets.stat_sum() - total[k]) < 1e-6
    for p, v in zip(plist, values):
        p.value = v


def test_batch_items_equal_to_vector_zero_are_shared(ctx):
    """In a scan most (tile, vector) items equal vector 0's and are evaluated once (dedup_*_kernel): the batch must
    be bit-identical to the per-vector calls, whichever vector carries the change, and when nothing changes."""
    from gammapy_b200 import Datasets, Map, configs
    from oracle import adapter

    ds = configs.make_c3(width=(8.0, 5.0), n_sources=6, n_reco=10, n_true=20)
    ds.counts = Map.from_geom(ds._geom, data=adapter.evaluator_for(ds).fake(21), dtype=float)
    datasets = Datasets([ds])
    values, plist = datasets.parameter_vector()
    names = [m.name for m in ds.models if m.name != ds.background_model.name]
    a, b = ds.models[names[0]], ds.models[names[3]]
    cols = {id(p): i for i, p in enumerate(plist)}
    grid = np.tile(values, (9, 1))
    grid[1, cols[id(a.spectral_model.amplitude)]] *= 1.3          # one source, spectral only
    grid[2, cols[id(b.spectral_model.amplitude)]] *= 0.0          # source switched off (amplitude == 0)
    grid[3, cols[id(b.spatial_model.lon_0)]] += 0.03              # one source, spatial
    grid[4, cols[id(ds.background_model.spectral_model.norm)]] *= 1.05   # every tile differs
    grid[5] = grid[1]                                             # same change twice: shares the flux vector
    grid[6, cols[id(a.spectral_model.amplitude)]] *= 1.3
    grid[6, cols[id(b.spatial_model.lon_0)]] += 0.03              # two changes at once
    # 7, 8: identical to vector 0
    batch = datasets.stat_sum_batch(grid)
    assert batch[5] == batch[1] and batch[7] == batch[0] and batch[8] == batch[0]
    assert len(set(batch[[0, 1, 2, 3, 4, 6]])) == 6
    for k in range(9):
        for p, v in zip(plist, grid[k]):
            p.value = v
        assert datasets.stat_sum() == batch[k], k
    for p, v in zip(plist, values):
        p.value = v
    # vector 0 is the odd one out: nothing may be shared with it
    grid2 = np.tile(grid[4], (3, 1))
    grid2[0] = values
    batch2 = datasets.stat_sum_batch(grid2)
    assert batch2[0] == batch[0] and batch2[1] == batch[4] and batch2[2] == batch[4]


def test_weighted_cash_dataset(ctx):
    """stat_type='cash_weighted' with a float mask (WeightedCashFitStatistic, fit_statistics.py:301-329)."""
    from gammapy_b200 import Map, MapDataset, configs
    from oracle import adapter, cash

    base = configs.make_c1(width=(1.6, 1.6), mask_radius=None)
    rng = np.random.RandomState(4)
    weights = rng.uniform(0, 1, size=base.data_shape)
    weights[rng.uniform(size=base.data_shape) < 0.3] = 0.0
    ds = MapDataset(models=base.models, counts=None, exposure=base.exposure, background=base.background,
                    psf=base.psf, edisp=base.edisp, mask_safe=None,
                    mask_fit=Map.from_geom(base._geom, data=weights, dtype=float), name=base.name,
                    stat_type="cash_weighted")
    de = adapter.evaluator_for(base, conv="exact64")
    counts = de.fake(6)
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    npred = de.npred()
    mask = ~(weights == False)  # noqa: E712
    ref = cash.reference_module()
    w32 = weights.astype(np.float32).astype(float)  # the device copy of the weights is float32
    if ref is not None:
        want = ref.weighted_cash_sum_cython(counts[mask], npred[mask], w32[mask])
    else:
        want = cash.weighted_cash_sum(counts[mask], npred[mask], w32[mask])
    got = ds.stat_sum_likelihood()
    assert abs(got - want) < 1e-3
    assert_allclose(ds.stat_array().sum(), got, rtol=1e-6)


def test_cached_work_descriptors_follow_the_state(ctx):
    """The per-tile overlap lists are kept between calls while sources, cutouts and slots are unchanged and rebuilt
    otherwise: a spectral step (kept), a position step (new cutout content, maybe new cutout), a per-model npred
    (other source list), a batched call (descriptors of another shape) and back -- every result against the oracle."""
    from gammapy_b200 import Datasets, Map, configs
    from oracle import adapter

    ds = configs.make_c2()
    de = adapter.evaluator_for(ds, conv="exact64")
    counts = de.fake(4)
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    datasets = Datasets([ds])

    def check(label):
        adapter.refresh_models(de, ds)
        got, want = datasets.stat_sum(), de.stat_sum()
        assert abs(got - want) < 1e-3, (label, got, want)
        return got

    base = check("first call")
    assert datasets.stat_sum() == base                           # descriptors reused
    ds.models["src-0"].spectral_model.index.value += 0.2
    ds.background_model.spectral_model.norm.value = 1.07
    check("spectral + background step")
    ds.models["src-1"].spatial_model.lon_0.value += 0.011      # inside the cutout margin: same cutout, new cube
    check("small move")
    ds.models["src-2"].spatial_model.lat_0.value += 0.9        # beyond r_eval + 0.1 deg: new cutout
    check("large move")
    part = ds.npred_signal(model_names=["src-3"]).data           # another source list on the same handle
    _, parts = de.npred_signal(per_model=True)
    assert_allclose(part, parts["src-3"], rtol=1e-5, atol=1e-7 * float(parts["src-3"].max()))
    check("after a per-model npred")
    values, plist = datasets.parameter_vector()
    grid = np.tile(values, (4, 1))
    grid[1:, [i for i, p in enumerate(plist) if p.name == "amplitude"][0]] *= [1.1, 1.2, 1.3]
    batch = datasets.stat_sum_batch(grid)
    assert abs(batch[0] - check("after a batched call")) < 1e-6
    ds.models["src-0"].spectral_model.amplitude.value = 0.0     # source drops out of the list
    check("amplitude zero")


def test_more_sources_than_a_descriptor_holds(ctx):
    """70 sources on one dataset: more than the 64 a work descriptor can list, but far fewer on any one tile, so the
    staged kernel still runs (the bound is per tile, from the cutout rows); npred and stat against the oracle."""
    from gammapy_b200 import configs

    ds = configs.make_c3(width=(8.0, 4.0), n_sources=70, n_reco=10, n_true=20)
    assert len(ds.models) == 71
    stat, want, npred, npred_want = _stat_pair(ds, seed=9)
    assert_allclose(npred, npred_want, rtol=1e-5, atol=1e-12 * npred_want.max())
    assert abs(stat - want) < 1e-3


def test_steady_state_replay_matches_the_oracle_and_the_batched_path(ctx):
    """A run of single-vector calls that only move spectral / background parameters is replayed from a CUDA graph
    (K0 reads the parameter vector on the device): every value equals the batched evaluation of the same vectors bit
    for bit and the oracle within tolerance; a spatial move in between falls back to the full path and back."""
    from gammapy_b200 import Datasets, Map, configs
    from oracle import adapter

    ds = configs.make_c2()
    de = adapter.evaluator_for(ds, conv="exact64")
    ds.counts = Map.from_geom(ds._geom, data=de.fake(7), dtype=float)
    datasets = Datasets([ds])
    values, plist = datasets.parameter_vector()
    names = [p.name for p in plist]
    i_index, i_norm, i_amp = names.index("index"), names.index("norm"), names.index("amplitude")
    grid = np.tile(values, (6, 1))
    grid[1:, i_index] += [0.05, -0.07, 0.11, 0.02, -0.03]
    grid[2:, i_norm] *= [1.02, 0.97, 1.05, 1.01]
    grid[4, i_amp] = 0.0           # a source drops out: not replayable, full path
    batch = datasets.stat_sum_batch(grid)
    launches = []
    for k in range(6):
        for p, v in zip(plist, grid[k]):
            p.value = v
        before = ctx.launch_count
        got = datasets.stat_sum()
        launches.append(ctx.launch_count - before)
        assert got == batch[k], k
        adapter.refresh_models(de, ds)
        assert abs(got - de.stat_sum()) < 1e-3, k
    assert launches[2] == 3 and launches[3] == 3   # K0 + fold + reduce, replayed
    plist[names.index("lon_0")].value += 0.013      # a position moves: spatial model + PSF stencil run again
    adapter.refresh_models(de, ds)
    assert abs(datasets.stat_sum() - de.stat_sum()) < 1e-3
    plist[i_index].value += 0.01
    adapter.refresh_models(de, ds)
    assert abs(datasets.stat_sum() - de.stat_sum()) < 1e-3
    for p, v in zip(plist, values):
        p.value = v


def test_residuals_excess_and_info_dict(ctx):
    """What one looks at after a fit (datasets/map.py:669-672, 1210-1255, 1791-1900): host arithmetic on the GPU's npred
    cubes, against the same arithmetic on the oracle's cubes."""
    from gammapy_b200 import Map, configs
    from oracle import adapter

    ds = configs.make_c2(width=(3.0, 3.0), mask_radius=1.0)
    de = adapter.evaluator_for(ds, conv="exact64")
    counts = de.fake(11)
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    npred, mask = de.npred(), np.asarray(ds.mask.data, bool)
    for method, fn in (("diff", lambda c, m: c - m), ("diff/model", lambda c, m: (c - m) / m),
                       ("diff/sqrt(model)", lambda c, m: (c - m) / np.sqrt(m))):
        got = ds.residuals(method).data
        assert np.isnan(got[~mask]).all()
        want = fn(counts[mask], npred[mask])
        assert_allclose(got[mask], want, rtol=2e-5, atol=2e-5 * np.abs(want).max())
    with pytest.raises(AttributeError):
        ds.residuals("ratio")
    with pytest.raises(NotImplementedError):
        ds.residuals(width="0.1 deg")
    assert_allclose(ds.excess.data, counts - np.asarray(ds.background.data, float))
    info = ds.info_dict()
    safe = np.asarray(ds.mask_safe.data, bool)
    assert info["name"] == ds.name and info["counts"] == int(counts[safe].sum())
    assert_allclose(info["npred"], npred[safe].sum(), rtol=1e-6)
    assert_allclose(info["npred_signal"], de.npred_signal()[safe].sum(), rtol=1e-6)
    assert_allclose(info["npred_background"] + info["npred_signal"], info["npred"], rtol=1e-9)
    assert_allclose(info["excess"], info["counts"] - info["npred_background"], rtol=1e-12)
    n, mu = info["counts"], info["npred_background"]
    assert_allclose(info["sqrt_ts"], np.sign(n - mu) * np.sqrt(2 * ((mu - n * np.log(mu)) - (n - n * np.log(n)))), rtol=1e-9)
    assert info["n_bins"] == counts.size and info["n_fit_bins"] == int(mask.sum()) and info["stat_type"] == "cash"
    assert_allclose(info["stat_sum"], ds.stat_sum(), rtol=0, atol=1e-9)
    assert 0 < info["exposure_min"] <= info["exposure_max"]
