"""Pin the oracle restatement against the reference's own data-free known-answer tests."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import cash, evaluator, spatial, spectral, wcs


# ---- stats/tests/test_fit_statistics.py:22-100, 117-150 ---------------------------------------------------
MU_SIG = [0.59752422, 9.13666449, 12.98288095, 5.56974565, 13.52509804, 11.81725635, 0.47963765, 11.17708176,
          5.18504894, 8.30202394]
N_ON = [0, 13, 7, 5, 11, 16, 0, 9, 3, 12]
CASH_REF = [1.19504844, -39.24635098872072, -9.925081055136996, -6.034002586236575, -30.249839537105466,
            -55.39143500383233, 0.9592753, -21.095413867175516, 0.49542219758430406, -34.19193611846045]


def test_cash_per_bin_golden():
    assert_allclose(cash.cash(N_ON, MU_SIG), CASH_REF)


def test_cash_sum_restatement_vs_per_bin():
    assert_allclose(cash.cash_sum(np.array(N_ON, float), np.array(MU_SIG)), np.sum(CASH_REF), rtol=1e-12)


def test_cash_sum_vs_compiled_reference():
    ref = cash.reference_module()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.RandomState(0)
    mu = rng.gamma(2.0, 2.0, 200000)
    mu[::1000] = 0.0
    mu[5::1000] = -1.0
    mu[7::1000] = 1e-26
    n = rng.poisson(np.clip(mu, 0, None)).astype(float)
    assert_allclose(cash.cash_sum(n, mu), ref.cash_sum_cython(n, mu), rtol=1e-13)
    w = rng.uniform(-0.2, 1.0, mu.size)
    assert_allclose(cash.weighted_cash_sum(n, mu, w), ref.weighted_cash_sum_cython(n, mu, w), rtol=1e-13)
    # truncation constant: Cython narrows 1e-25 to float before the log
    z = np.zeros(3)
    assert ref.cash_sum_cython(np.ones(3), z) == cash.cash_sum(np.ones(3), z)
    assert cash.cash_sum(np.ones(3), z) == pytest.approx(2 * 3 * (1e-25 - cash.LOG_TRUNCATION), rel=1e-15)


def test_cash_empty():
    assert cash.cash_sum(np.zeros(0), np.zeros(0)) == 0.0


# ---- modeling/models/tests/test_spectral.py:59-190, utils/tests/test_integrate.py:8-16 ---------------------
def test_spectral_goldens():
    pl = {"type": "pl", "index": 2.3, "amplitude": 4.0, "reference": 1.0}
    assert_allclose(spectral.evaluate(pl, 2.0), 4 * 2.0 ** (-2.3), rtol=1e-7)
    assert_allclose(spectral.integral(pl, 1.0, 10.0), 2.9227116204223784, rtol=1e-7)
    pl2 = {"type": "pl", "index": 2.0, "amplitude": 4.0, "reference": 1.0}
    assert_allclose(spectral.integral(pl2, 1.0, 10.0), 3.6, rtol=1e-7)
    ecpl = {"type": "ecpl", "index": 1.6, "amplitude": 4.0, "reference": 1.0, "lambda_": 0.1, "alpha": 1.0}
    assert_allclose(spectral.evaluate(ecpl, 2.0), 1.080321705479446, rtol=1e-7)
    assert_allclose(spectral.integral(ecpl, [1.0], [10.0]), 3.765838739678921, rtol=1e-5)
    lp = {"type": "lp", "alpha": 2.3, "amplitude": 4.0, "reference": 1.0, "beta": 0.5}
    assert_allclose(spectral.evaluate(lp, 2.0), 0.6387956571420305, rtol=1e-7)
    assert_allclose(spectral.integral(lp, [1.0], [10.0]), 2.255689748270628, rtol=1e-5)
    norm = {"type": "pl-norm", "tilt": 2.0, "norm": 4.0, "reference": 1.0}
    assert_allclose(spectral.evaluate(norm, 2.0), 1.0)


def test_trapz_loglog_powerlaw_exact():
    pl = {"type": "pl", "index": 2.3, "amplitude": 1e-12, "reference": 1.0}
    energy = np.array([1.0, 10.0])
    val = spectral.trapz_loglog(spectral.evaluate(pl, energy), energy)
    assert_allclose(val, spectral.integral(pl, 1.0, 10.0), rtol=1e-12)


def test_pl_index_one_uses_log_branch():
    pl = {"type": "pl", "index": 1.0, "amplitude": 2.0, "reference": 1.0}
    assert_allclose(spectral.integral(pl, np.array([1.0]), np.array([10.0])), 2.0 * np.log(10.0))


# ---- modeling/models/tests/test_spatial.py:61-245, 659-691 ------------------------------------------------
def test_sky_gaussian_ratio():
    v0 = spatial.gauss_evaluate(5.0, 15.0, 5.0, 15.0, 1.0)
    vs = spatial.gauss_evaluate(5.0, 16.0, 5.0, 15.0, 1.0)
    assert_allclose(v0 / vs, np.exp(0.5))


def test_sky_gaussian_elongated():
    sigma, semi_minor = 4.0, 2.0
    e = np.sqrt(1 - (semi_minor / sigma) ** 2)
    v0 = spatial.gauss_evaluate(0.0, 0.0, 0.0, 0.0, sigma, e, 0.0)
    assert_allclose(v0 / spatial.gauss_evaluate(0.0, 4.0, 0.0, 0.0, sigma, e, 0.0), np.exp(0.5))
    assert_allclose(v0 / spatial.gauss_evaluate(2.0, 0.0, 0.0, 0.0, sigma, e, 0.0), np.exp(0.5))
    assert_allclose(v0 / spatial.gauss_evaluate(0.0, 2.0, 0.0, 0.0, sigma, e, 90.0), np.exp(0.5))


def test_sky_disk():
    val = spatial.disk_evaluate(np.array([1.0, 5.0, 359.0]), 46.0, 1.0, 45.0, 2.0)
    assert_allclose(val, [261.263956, 0, 261.263956], rtol=1e-7, atol=1e-12)
    m = {"type": "disk", "lon_0": 1.0, "lat_0": 45.0, "r_0": 2.0}
    assert_allclose(spatial.evaluation_radius(m), 2.222)
    assert_allclose(spatial.evaluation_bin_size_min(m), 0.198)
    assert_allclose(spatial.disk_evaluate(0.0, 1.5, 0.0, 0.0, 2.0, np.sqrt(0.75), 90.0), 0, atol=1e-12)
    assert_allclose(spatial.disk_evaluate(0.0, 90.0, 0.0, 90.0, 5.0) * 2 * np.pi * (1 - np.cos(np.deg2rad(5.0))), 1)


def test_sky_disk_edge():
    def d(lon, lat):
        return spatial.disk_evaluate(lon, lat, 0.0, 0.0, 2.0, e=0.5, phi=0.0)

    assert_allclose(d(0, 2.0) / d(0, 0), 0.5)
    assert_allclose(d(0, 2.0 + 0.01) / d(0, 0), 0.05)
    assert_allclose(d(0, 2.0 - 0.01) / d(0, 0), 0.95)


@pytest.mark.parametrize("sigma", [1e-4, 1e-3, 1e-2, 1e-1])
def test_integrate_wcs_geom_sum_one(sigma):
    geom = wcs.ImageGeom.create(skydir=(0, 0), binsz=0.02, width=(2, 2))
    w = spatial.integrate_geom({"type": "gauss", "lon_0": 0.0, "lat_0": 0.0, "sigma": sigma}, geom)
    assert_allclose(w.sum(), 1, rtol=1e-3)


def test_integrate_geom_no_overlap():
    geom = wcs.ImageGeom.create(skydir=(0, 0), binsz=0.02, width=(2, 2))
    w = spatial.integrate_geom({"type": "gauss", "lon_0": 50.0, "lat_0": 0.0, "sigma": 0.1}, geom)
    assert w.shape == (100, 100)
    assert_allclose(w.sum(), 0, atol=1e-30)


def test_point_source_weights():
    geom = wcs.ImageGeom.create(skydir=(0, 0), binsz=1.0, npix=(4, 4))
    w = spatial.integrate_geom({"type": "point", "lon_0": 0.25, "lat_0": -0.3}, geom)
    assert_allclose(w.sum(), 1)
    assert np.count_nonzero(w) == 4
    x0, y0 = geom.world_to_pix(0.25, -0.3)
    yy, xx = np.nonzero(w)
    assert_allclose(np.sum(w[yy, xx] * xx), x0)  # centroid preserved
    assert_allclose(np.sum(w[yy, xx] * yy), y0)


# ---- modeling/models/tests/test_cube.py:802-845 -----------------------------------------------------------
def test_evaluate_integrate_nd_geom_golden():
    geom = wcs.ImageGeom.create(skydir=(0, 0), binsz=0.05, width=(1, 1.2))
    assert geom.data_shape == (24, 20)
    sm = {"type": "gauss", "lon_0": 0.0, "lat_0": 0.0, "sigma": 0.1}
    pl = {"type": "pl", "index": 2.0, "amplitude": 1e-11, "reference": 1.0}
    edges = np.geomspace(1, 10, 4)
    flux = spectral.integral(pl, edges[:-1], edges[1:])
    w = spatial.integrate_geom(sm, geom)
    assert w.dtype == np.float64
    assert_allclose(flux * w[12, 10], [1.973745e-13, 9.161312e-14, 4.252304e-14], rtol=1e-6)
    lon, lat = geom.get_coord()
    centers = np.exp(0.5 * (np.log(edges[:-1]) + np.log(edges[1:])))
    dnde = spatial.evaluate(sm, lon, lat)[12, 10] * spectral.evaluate(pl, centers)
    assert_allclose(dnde, [2.278184e-07, 4.908198e-08, 1.057439e-08], rtol=1e-6)


# ---- maps/wcs/tests/test_geom.py:279-329 (Cutout2D index rule) ---------------------------------------------
def test_cutout_rules():
    geom = wcs.ImageGeom.create(skydir=(0, 0), npix=10, binsz=0.5)
    cut, (ys, xs) = geom.cutout_info((0.0, 0.0), 2.0)
    assert cut.data_shape == (4, 4)
    assert (ys.start, xs.start) == (3, 3)
    geom = wcs.ImageGeom.create(skydir=(0, 0), npix=10, binsz=0.1)
    cut, (ys, xs) = geom.cutout_info((0.1, 0.2), 0.6)
    assert cut.data_shape == (6, 6)
    c = cut.center_skydir
    assert_allclose(((c[0] + 180) % 360 - 180, c[1]), (0.1, 0.2), atol=1e-12)
    cut, _ = geom.cutout_info((0.1, 0.2), 0.01)  # below one pixel: min_npix = 1
    assert cut.data_shape == (1, 1)
    cut, _ = geom.cutout_info((0.1, 0.2), 0.62, odd_npix=True)
    assert cut.data_shape == (7, 7)


def test_solid_angle_small_pixel():
    geom = wcs.ImageGeom.create(skydir=(0, 0), binsz=0.02, width=(2, 2))
    omega = geom.solid_angle()
    assert_allclose(omega[50, 50], np.deg2rad(0.02) ** 2, rtol=1e-4)
    assert_allclose(omega.sum(), 4 * np.deg2rad(1) * np.sin(np.deg2rad(1)), rtol=1e-4)


# ---- datasets/tests/test_map.py:1418-1445 --------------------------------------------------------------------
def test_npred_psf_after_edisp_total():
    from gammapy_b200 import (FoVBackgroundModel, MapAxis, MapDataset, PointSpatialModel, PowerLawSpectralModel,
                              PSFMap, SkyModel, WcsGeom)
    from oracle import adapter

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
    npred = adapter.evaluator_for(dataset).npred()
    assert_allclose(npred.sum(), 129553.858658)


# ---- datasets/tests/test_evaluator.py semantics ---------------------------------------------------------------
def test_evaluator_cache_and_update_rules():
    from gammapy_b200 import configs
    from oracle import adapter

    ds = configs.make_c1(width=(2.0, 2.0), mask_radius=0.9)
    de = adapter.evaluator_for(ds)
    ev = de.evaluators["model-simu"]
    de.npred()
    first = ev.slices
    flux = ev._flux_spatial
    # spectral change: spatial flux cached (evaluator.py:282-286)
    ev.model["spectral"]["index"] = 2.5
    de.npred()
    assert ev._flux_spatial is flux
    # small move: recomputed on the same cutout (evaluator.py:465-481)
    ev.model["spatial"]["lon_0"] = 0.25
    de.npred()
    assert ev._flux_spatial is not flux and ev.slices == first
    # amplitude zero -> zeros
    ev.model["spectral"]["amplitude"] = 0.0
    assert_allclose(de.npred_signal().sum(), 0)


def test_oracle_noise_floor_small():
    """The reference's float32 FFT agrees with exact arithmetic to ~1e-6 relative on total npred."""
    from gammapy_b200 import configs
    from oracle import adapter

    ds = configs.make_c1(width=(2.0, 2.0), mask_radius=0.9)
    a = adapter.evaluator_for(ds, conv="fft32").npred()
    b = adapter.evaluator_for(ds, conv="exact64").npred()
    assert np.max(np.abs(a - b) / b) < 5e-6


def test_golden_fixture_file():
    """tests/golden/reference_kats.json (transcribed reference constants) against the oracle."""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")
    gold = json.load(open(path))
    c = gold["cash"]
    assert_allclose(cash.cash(c["n_on"], c["mu_sig"]), c["cash"])
    s = gold["spectral"]
    pl = dict(s["powerlaw"], type="pl")
    assert_allclose(spectral.integral(pl, 1.0, 10.0), pl["integral_1_10TeV"], rtol=1e-7)
    ecpl = dict(s["ecpl"], type="ecpl")
    assert_allclose(spectral.integral(ecpl, [1.0], [10.0]), ecpl["integral_1_10TeV"], rtol=1e-5)
    lp = dict(s["logpar"], type="lp")
    assert_allclose(spectral.integral(lp, [1.0], [10.0]), lp["integral_1_10TeV"], rtol=1e-5)
    assert_allclose(spatial.disk_evaluate(1.0, 46.0, 1.0, 45.0, 2.0), gold["spatial"]["disk_value"], rtol=1e-7)


def test_psf_kernel_from_disk_models_to_image_golden():
    """irf/psf/tests/test_kernel.py:39-67 ``test_psf_kernel_to_image``: kernels built from two disk models
    (``PSFKernel.from_spatial_model``, irf/psf/kernel.py:99-126: odd-sized geometry, 4 x upsampled ``integrate_geom`` with
    oversampling factor 1, block sum) and weighted over energy with a power law of index 2 (``to_image``, :165-212).
    Pins the oracle's disk evaluation with its smoothed edge, the upsample / block-sum chain and the solid angles at
    the reference's own 1e-5."""
    from oracle.wcs import ImageGeom

    geom = ImageGeom.create(skydir=(0.0, 0.0), binsz=0.1, npix=(50, 50)).to_odd_npix(max_radius=2.5)
    assert (geom.nx, geom.ny) == (51, 51)
    up = geom.upsample(4)
    planes = []
    for r_0 in (0.5, 0.2):
        model = {"type": "disk", "lon_0": geom.center_skydir[0], "lat_0": geom.center_skydir[1], "r_0": r_0, "e": 0.0,
                 "phi": 0.0, "edge_width": 0.01}
        fine = spatial.integrate_geom(model, up, oversampling_factor=1)
        planes.append(fine.reshape(geom.ny, 4, geom.nx, 4).sum(axis=(1, 3)))
    planes = np.array(planes)
    planes /= planes.sum(axis=(1, 2), keepdims=True)  # PSFKernel normalises every plane (irf/psf/kernel.py:70-80)
    edges = np.geomspace(1.0, 10.0, 3)
    flux = 1 / edges[:-1] - 1 / edges[1:]  # integral of E^-2 over each bin
    for exposure, want in (([1.0, 1.0], (0.028096, 0.009615)), ([1.0, 2.0], (0.037555, 0.007752))):
        w = flux * np.array(exposure)
        image = (planes * (w / w.sum())[:, None, None]).sum(axis=0)
        assert_allclose(image.sum(), 1.0, atol=1e-5)
        assert_allclose((image[25, 25], image[22, 22], image[20, 20]), want + (0.0,), atol=1e-5)
