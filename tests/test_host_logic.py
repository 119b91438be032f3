"""Host-side mirror of the gammapy API: axes, geometry, parameters, models, static IRF builders (CPU only)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from gammapy_b200 import (DiskSpatialModel, EDispKernelMap, FoVBackgroundModel, GaussianSpatialModel, Map, MapAxis,
                          MapDataset, Models, Parameter, PointSpatialModel, PowerLawSpectralModel, PSFMap, SkyModel,
                          WcsGeom, configs)
from oracle import adapter, wcs


def test_map_axis_energy():
    ax = MapAxis.from_energy_bounds("0.1 TeV", "10 TeV", nbin=10)
    assert ax.nbin == 10 and ax.name == "energy"
    assert_allclose(ax.edges, np.geomspace(0.1, 10, 11), rtol=1e-14)
    assert_allclose(ax.center, np.sqrt(ax.edges[:-1] * ax.edges[1:]), rtol=1e-14)
    ax2 = MapAxis.from_energy_bounds(100, 1e4, nbin=2, unit="GeV", name="energy_true")
    assert_allclose(ax2.edges, [0.1, 1.0, 10.0], rtol=1e-14)
    assert ax.copy(name="energy_true").name == "energy_true" and ax.name == "energy"
    rad = MapAxis.from_bounds(0, 0.66, nbin=66, name="rad", unit="deg")
    assert_allclose(rad.center[0], 0.005)
    assert rad.upsample(20).nbin == 1320


def test_wcs_geom_matches_oracle_geometry():
    geom = WcsGeom.create(skydir=(0, 0), binsz=0.02, width=(4, 3), frame="galactic")
    og = adapter.geom_to_oracle(geom)
    assert geom.data_shape == (150, 200) == og.data_shape
    assert_allclose(geom.solid_angle(), og.solid_angle(), rtol=1e-14)
    lon, lat = og.get_coord()
    c = geom.get_coord()
    assert_allclose(c["lon"], lon)
    assert_allclose(c["lat"], lat)
    ys, xs = geom.cutout_slices_for((0.2, 0.1), 4.52, odd_npix=True)
    _, (oys, oxs) = og.cutout_info((0.2, 0.1), 4.52, odd_npix=True)
    assert (ys, xs) == (oys, oxs)
    assert geom.to_odd_npix(max_radius=0.66).npix == (67, 67)
    assert WcsGeom.create(skydir=(10, 5), binsz=0.1, width=1).crval == (10.0, 5.0)  # oblique CAR, as the reference builds it
    assert WcsGeom.create(skydir=(10, 5), binsz=0.1, width=1, proj="TAN").crval == (10.0, 5.0)
    with pytest.raises(NotImplementedError):
        WcsGeom.create(binsz=0.1, width=1, proj="AIT")


def test_parameter_factor_scale_contract():
    """modeling/parameter.py:524-574: scale10 autoscale, factor <-> value."""
    par = Parameter("amplitude", "3e-12 cm-2 s-1 TeV-1")
    assert par.unit == "cm-2 s-1 TeV-1" and par.value == 3e-12
    par.autoscale()
    assert_allclose(par.scale, 1e-12)
    assert_allclose(par.factor, 3.0)
    par.factor = 4.5
    assert_allclose(par.value, 4.5e-12)
    lim = Parameter("index", 2.0, min=1, max=5)
    assert_allclose([lim.factor_min, lim.factor_max], [1, 5])
    scan = Parameter("norm", 1.0, error=0.1)
    assert_allclose(scan.scan_values, np.linspace(0.8, 1.2, 11))
    with pytest.raises(TypeError):
        Parameter("x", 1.0, frozen="yes")


def test_models_parameters_and_selection():
    spatial = GaussianSpatialModel(lon_0="0.2 deg", lat_0="0.1 deg", sigma="0.3 deg", frame="galactic")
    spectral = PowerLawSpectralModel(index=3, amplitude="1e-11 cm-2 s-1 TeV-1", reference="1 TeV")
    model = SkyModel(spatial_model=spatial, spectral_model=spectral, name="src")
    assert model.parameters.names == ["lon_0", "lat_0", "sigma", "e", "phi", "index", "amplitude", "reference"]
    assert [p.name for p in model.parameters.free_parameters] == ["lon_0", "lat_0", "sigma", "index", "amplitude"]
    assert_allclose(model.evaluation_radius, 1.5)
    disk = DiskSpatialModel(r_0=2.0)
    assert_allclose([disk.evaluation_radius, disk.evaluation_bin_size_min], [2.222, 0.198])
    assert PointSpatialModel().evaluation_radius == 0
    bkg = FoVBackgroundModel(dataset_name="a")
    models = Models([model, bkg])
    assert models.select(datasets_names="a").names == ["src", "a-bkg"]
    assert models.select(datasets_names="b").names == ["src"]
    with pytest.raises(TypeError):
        model.spectral_model.index = 2.0  # parameters are assigned through .value, as in the reference
    # independent copies per instance
    other = PowerLawSpectralModel()
    other.index.value = 1.0
    assert spectral.index.value == 3


def test_static_irf_builders():
    e_true = MapAxis.from_energy_bounds(0.03, 30, 20, name="energy_true")
    e_reco = MapAxis.from_energy_bounds(0.1, 10, 10)
    geom = WcsGeom.create(skydir=(0, 0), binsz=0.02, width=4, frame="galactic", axes=[e_true])
    sigma = np.geomspace(0.12, 0.04, 20)
    kernel = PSFMap.from_gauss(e_true, sigma=sigma).get_psf_kernel(geom, containment=0.999)
    data = kernel.data
    assert data.shape[0] == 20 and data.shape[1] == data.shape[2] and data.shape[1] % 2 == 1
    assert_allclose(data.sum(axis=(1, 2)), 1.0, rtol=1e-12)
    assert np.all(data >= 0)
    h = data.shape[1] // 2
    assert np.all(data[:, h, h] == data.max(axis=(1, 2)))  # peaked at the centre pixel
    # narrower PSF at high energy -> smaller support
    support = [(np.abs(np.nonzero(p)[0] - h)).max() for p in data]
    assert support[0] > support[-1]
    # containment radius of a Gaussian: r = sigma sqrt(-2 ln(1 - f)), grid resolution 0.0005 deg
    psf = PSFMap.from_gauss(e_true, sigma=0.05)
    assert_allclose(psf.containment_radius(0.68), 0.05 * np.sqrt(-2 * np.log(0.32)), atol=6e-4)
    edisp = EDispKernelMap.from_gauss(e_reco, e_true, sigma=0.15, bias=0).get_edisp_kernel().pdf_matrix
    assert edisp.shape == (20, 10)
    assert np.all(edisp.sum(axis=1) <= 1 + 1e-12) and edisp.sum(axis=1).max() > 0.99
    diag = EDispKernelMap.from_diagonal_response(e_reco, e_true).get_edisp_kernel().pdf_matrix
    assert_allclose(diag.sum(axis=1)[(e_true.edges[:-1] >= 0.1) & (e_true.edges[1:] <= 10)], 1.0)


def test_map_dataset_create_and_masks():
    ds = configs.make_c1(width=(1.0, 1.0), mask_radius=0.4)
    assert ds.counts.data.dtype == np.float32 and ds.exposure.data.dtype == np.float32
    assert ds.mask.data.dtype == bool
    assert ds.mask_image.data.shape == (50, 50)
    assert ds.mask_image.data.sum() == (ds.mask_fit.data[0]).sum()
    assert ds.models.names == ["model-simu", "c1-bkg"]
    empty = MapDataset.create(ds._geom)
    assert not empty.mask_safe.data.any()  # as the reference: all-False safe mask until the user sets it
    m = Map.from_geom(ds._geom)
    assert m.data.dtype == np.float32  # maps/core.py:269 default dtype
    assert (m + 1.5).data.dtype == np.float64 or (m + 1.5).data.dtype == np.float32


def test_oracle_adapter_roundtrip():
    ds = configs.make_c2(width=(2.0, 2.0), mask_radius=0.9)
    de = adapter.evaluator_for(ds)
    assert len(de.evaluators) == 5 and de.bkg_model is not None
    assert de.ds.exposure.shape == (20, 100, 100) and de.ds.edisp.shape == (20, 10)
    assert_allclose(de.ds.energy_center, ds._geom.axes["energy"].center)


def test_ts_map_host_logic():
    """Argument checks and the default position mask of the TS-map front end run before any device call."""
    import pytest

    from gammapy_b200.estimators import BrentqFluxEstimator, TSMapEstimator, compute_ts_map

    counts = np.zeros((2, 5, 6))
    with pytest.raises(ValueError):
        compute_ts_map(counts, counts[:1], counts, np.ones((2, 3, 3)))          # shapes differ
    with pytest.raises(ValueError):
        compute_ts_map(counts, counts, counts, np.ones((1, 3, 3)))              # one kernel plane per map plane
    with pytest.raises(ValueError):
        compute_ts_map(np.zeros(5), np.zeros(5), np.zeros(5), np.ones((1, 3, 3)))
    with pytest.raises(NotImplementedError):
        BrentqFluxEstimator(rtol=0.01, n_sigma=1, selection_optional=["ul"])
    with pytest.raises(NotImplementedError):
        BrentqFluxEstimator(rtol=0.01, n_sigma=1, ts_threshold=4.0)
    # estimators/map/ts.py:500-505: any plane selected and a positive summed background
    mask = np.zeros((2, 5, 6), dtype=bool)
    mask[1, 2:4, 1:5] = True
    background = np.ones((2, 5, 6))
    background[:, 3, 4] = 0.0
    got = TSMapEstimator._estimate_mask_2d({"mask": mask, "background": background})
    want = np.zeros((5, 6), dtype=bool)
    want[2:4, 1:5] = True
    want[3, 4] = False
    assert np.array_equal(got, want)
    with pytest.raises(ValueError, match="No valid positions"):
        TSMapEstimator().estimate_flux_map({"mask": np.zeros((1, 5, 6), dtype=bool), "background": background[:1],
                                            "counts": counts[:1], "exposure": counts[:1], "kernel": np.ones((1, 3, 3))})


def test_fit_with_the_minimum_on_a_parameter_limit():
    """a fit whose unconstrained minimum lies outside a limit must end ON the limit with success, the
    other parameters at the constrained optimum, and Hesse must not evaluate points outside the limits."""
    from gammapy_b200.distributed import ShardedDatasets
    from gammapy_b200.modeling import Fit
    from gammapy_b200.modeling.models import PowerLawSpectralModel, PointSpatialModel, SkyModel

    model = SkyModel(spatial_model=PointSpatialModel(lon_0=0.0, lat_0=0.0, frame="galactic"),
                     spectral_model=PowerLawSpectralModel(index=2.0, amplitude=1e-12, reference=1), name="src")
    for name in ("lon_0", "lat_0"):
        getattr(model.spatial_model, name).frozen = True
    index, amplitude = model.spectral_model.index, model.spectral_model.amplitude
    index.min, index.max = 2.5, 4.0      # the unconstrained minimum (index = 2.0) lies below the limit
    index.value = 3.0
    pars = list(model.parameters)
    names = [p.name for p in pars]
    seen = []

    def local_stat(values):
        seen.append(values[:, names.index("index")].copy())
        i, a = values[:, names.index("index")], np.log(values[:, names.index("amplitude")] / 2e-12)
        return 40 * (i - 2.0) ** 2 + 30 * a ** 2 + 24 * (i - 2.0) * a

    sharded = ShardedDatasets(None, local_stat=local_stat, global_models=[model], local_parameters=pars)
    result = Fit().run(sharded)
    assert result.success, result.message
    assert index.value == pytest.approx(2.5, abs=1e-9)
    # constrained optimum in log-amplitude: 60 a + 24 (2.5 - 2.0) = 0
    # (the EDM goal 2e-4 stops within 30 d^2 < 2e-4, i.e. |d| < 2.6e-3 = 0.014 sigma of the optimum)
    assert np.log(amplitude.value / 2e-12) == pytest.approx(-0.2, abs=3e-3)
    assert min(float(v.min()) for v in seen) >= 2.5 - 1e-12
    assert max(float(v.max()) for v in seen) <= 4.0 + 1e-12
    assert np.isfinite(amplitude.error) and amplitude.error > 0


def test_units_are_converted_or_refused_and_boolean_masks_select():
    from gammapy_b200.modeling.parameter import Parameter, Parameters
    from gammapy_b200.utils import parse_quantity

    assert parse_quantity("1e-9 cm-2 s-1 GeV-1") == (pytest.approx(1e-6), "cm-2 s-1 TeV-1")
    assert parse_quantity("2 m-2 s-1 TeV-1") == (pytest.approx(2e-4), "cm-2 s-1 TeV-1")
    assert parse_quantity("0.1 GeV-1") == (pytest.approx(100.0), "TeV-1")
    assert parse_quantity("3 cm2 s") == (pytest.approx(3e-4), "m2 s")
    with pytest.raises(ValueError):
        parse_quantity("1 erg-1 cm-2 s-1")
    pars = Parameters([Parameter("a", 1.0), Parameter("b", 2.0), Parameter("c", 3.0)])
    assert pars[np.array([True, False, True])].names == ["a", "c"]
    assert pars[[0, 2]].names == ["a", "c"]
    with pytest.raises(IndexError):
        pars[np.array([True, False])]


def test_parameter_lazy_factor_and_engine_mirror():
    """The value setter leaves the factor to be computed when first read (same scale as the reference's eager
    computation, modeling/parameter.py:229-233) and keeps the vectors of the engines the parameter is bound to equal to
    the value; copies and pickles of a bound parameter are bound to nobody."""
    import copy
    import pickle

    from gammapy_b200.modeling.parameter import bind_mirror, release_mirror

    p = Parameter("amplitude", "3e-12 cm-2 s-1 TeV-1")
    q = Parameter("index", 2.0)
    vec = bind_mirror([p, q])
    other = bind_mirror([q])
    assert vec.tolist() == [3e-12, 2.0] and other.tolist() == [2.0]
    p.value = 5e-12
    q.value = 2.5
    assert vec.tolist() == [5e-12, 2.5] and other.tolist() == [2.5]
    assert p.factor == 5e-12 and p.scale == 1.0
    p.autoscale()  # scale10: factor in [1, 10), value unchanged
    assert p.scale == 1e-12 and p.factor == pytest.approx(5.0) and p.value == 5e-12 == vec[0]
    p.value = 7e-12  # factor follows with the scale it has now
    assert p.factor == pytest.approx(7.0)
    p.factor = 2.0
    assert p.value == pytest.approx(2e-12) and vec[0] == p.value
    p.reset_autoscale()
    assert p.scale == 1.0 and p.factor == p.value
    for clone in (copy.copy(p), copy.deepcopy(p), pickle.loads(pickle.dumps(p))):
        assert clone._mirrors == () and clone.value == p.value and clone.factor == p.factor
        clone.value = 1.0
        assert vec[0] == p.value != 1.0
    release_mirror([p, q], vec)
    q.value = 3.0
    assert vec[1] == 2.5 and other[0] == 3.0 and p._mirrors == ()


def test_energy_flux_against_the_reference_goldens():
    """``SpectralModel.energy_flux`` on the host (spectral.py:423-445; analytic for the power laws :1070-1100,
    :1182-1200, the log-log trapezoid otherwise): ``eflux_1_10TeV`` and ``val_at_2TeV`` of the reference's
    ``TEST_MODELS`` (modeling/models/tests/test_spectral.py:59-190)."""
    from gammapy_b200.modeling.models import (ExpCutoffPowerLawSpectralModel, LogParabolaSpectralModel,
                                              PowerLawNormSpectralModel, PowerLawSpectralModel)

    cases = [
        (PowerLawSpectralModel(index=2.3, amplitude=4, reference=1), 4 * 2.0 ** (-2.3), 6.650836884969039),
        (PowerLawSpectralModel(index=2, amplitude=4, reference=1), 1.0, 9.210340371976184),
        (PowerLawNormSpectralModel(tilt=2, norm=4, reference=1), 1.0, 9.210340371976184),
        (ExpCutoffPowerLawSpectralModel(index=1.6, amplitude=4, reference=1, lambda_=0.1), 1.080321705479446, 9.901735870666526),
        (LogParabolaSpectralModel(alpha=2.3, amplitude=4, reference=1, beta=0.5), 0.6387956571420305, 3.9586515834989267),
    ]
    for model, val_at_2, eflux in cases:
        assert_allclose(model(2.0), val_at_2, rtol=1e-12)
        assert_allclose(model.energy_flux(1.0, 10.0), eflux, rtol=1e-7)  # the reference's own tolerance
    lp = cases[-1][0]
    parts = lp.energy_flux(np.array([1.0, 3.0]), np.array([3.0, 10.0]))
    assert parts.shape == (2,) and abs(parts.sum() - lp.energy_flux(1.0, 10.0)) < 1e-4 * parts.sum()


def test_cash_counts_statistic_goldens():
    """``cash_sqrt_ts`` (what ``MapDataset.info_dict`` reports) against the reference's ``test_cash_basic`` values
    (stats/tests/test_counts_statistic.py:9-29): (n_on, mu_bkg) -> excess, sqrt(TS)."""
    from gammapy_b200.stats import cash, cash_sqrt_ts

    for n_on, mu_bkg, excess, sqrt_ts in [(1, 2, -1.0, -0.78339367), (5, 1, 4.0, 2.84506224), (10, 5, 5.0, 1.96543726),
                                          (100, 23, 77.0, 11.8294207), (1, 20, -19, -5.65760863)]:
        assert n_on - mu_bkg == excess
        assert_allclose(cash_sqrt_ts(n_on, mu_bkg), sqrt_ts, atol=1e-5)
    grid = cash_sqrt_ts(5 * np.ones((3, 2, 4)), np.ones((3, 2, 4)))
    assert grid.shape == (3, 2, 4) and np.allclose(grid, 2.84506224, atol=1e-5)
    assert cash_sqrt_ts(0, 0) == 0.0                       # no counts, no background: mu truncated, TS = 0
    assert_allclose(cash(0, 0.0), 2e-25)
    assert_allclose(cash(3, 2.0), 2 * (2.0 - 3 * np.log(2.0)))


def test_edisp_kernel_builders_against_the_reference_goldens():
    """``EDispKernel.from_gauss`` / ``from_diagonal_response`` against the reference's data-free values
    (irf/edisp/tests/test_kernel.py:26-70): the reco-summed response of a Gaussian kernel with sigma 0.2, bias 0.1
    (rtol 1e-3, the reference's own tolerance -- it integrates an interpolated migra grid, here the error function is
    integrated directly) and the 0 / 1 pattern of the diagonal response between different axes."""
    from gammapy_b200 import EDispKernel, MapAxis

    energy_axis = MapAxis.from_energy_bounds("0.1 TeV", "10 TeV", nbin=3)
    energy_axis_true = MapAxis.from_energy_bounds("0.08 TeV", "20 TeV", nbin=5, name="energy_true")
    edisp = EDispKernel.from_gauss(energy_axis_true, energy_axis, sigma=0.2, bias=0.1)
    assert edisp.pdf_matrix.shape == (5, 3)
    assert_allclose(edisp.pdf_matrix.sum(axis=1), [0.97142, 1.0, 1.0, 1.0, 0.12349], rtol=1e-3)
    e_true = MapAxis.from_energy_edges([0.5, 1, 2, 4, 6], name="energy_true")
    diag = EDispKernel.from_diagonal_response(e_true, MapAxis.from_energy_edges([2, 4, 6]))
    assert diag.pdf_matrix.shape == (4, 2)
    np.testing.assert_array_equal(diag.pdf_matrix, [[0, 0], [0, 0], [1, 0], [0, 1]])
    square = EDispKernel.from_diagonal_response(e_true, e_true.copy(name="energy"))
    assert square.pdf_matrix[0][0] == 1 and square.pdf_matrix[2][0] == 0 and square.pdf_matrix.sum() == 4
