"""Model serialisation in the reference's YAML layout (modeling/models/core.py:545-700, cube.py:463-522, 784-835,
parameter.py:503-522; gammapy_b200/modeling/io.py): a file in the layout of the reference's own example
(modeling/models/tests/data/examples.yaml: per-parameter ``scale`` keys, ``.nan`` limits) is read, models written here
are read back unchanged -- values, units, limits, frozen flags, frames, priors, shared parameters -- and give the same
npred in the oracle."""
import numpy as np
import pytest

from gammapy_b200 import Models, configs
from gammapy_b200.modeling import GaussianPrior, UniformPrior
from gammapy_b200.modeling.models import (ExpCutoffPowerLawSpectralModel, FoVBackgroundModel, GaussianSpatialModel,
                                          PointSpatialModel, SkyModel)

REFERENCE_LAYOUT = """
components:
- name: source0
  type: SkyModel
  spatial:
    type: PointSpatialModel
    frame: galactic
    parameters:
    - name: lon_0
      value: -0.5
      scale: 0.01
      unit: deg
      min: -180.0
      max: 180.0
      frozen: true
    - name: lat_0
      value: -0.0005
      scale: 0.01
      unit: deg
      min: -90.0
      max: 90.0
      frozen: true
  spectral:
    type: ExpCutoffPowerLawSpectralModel
    parameters:
    - name: index
      value: 2.1
      scale: 1.0
      unit: ''
      min: .nan
      max: .nan
      frozen: false
    - name: amplitude
      value: 2.3e-12
      scale: 1.0e-12
      unit: cm-2 s-1 TeV-1
      min: .nan
      max: .nan
      frozen: false
    - name: reference
      value: 1.0
      scale: 1.0
      unit: TeV
      min: .nan
      max: .nan
      frozen: true
    - name: lambda_
      value: 0.006
      scale: 0.1
      unit: TeV-1
      min: .nan
      max: .nan
      frozen: false
- name: source1
  type: SkyModel
  datasets_names: [obs-1]
  apply_irf: {exposure: true, psf: false, edisp: true}
  spatial:
    type: gauss
    frame: galactic
    parameters:
    - {name: lon_0, value: 0.3, unit: deg}
    - {name: lat_0, value: 0.1, unit: deg}
    - {name: sigma, value: 0.2, unit: deg, min: 0.0}
  spectral:
    type: pl
    parameters:
    - {name: index, value: 2.4, error: 0.1}
    - {name: amplitude, value: 1.0e-12, unit: cm-2 s-1 TeV-1}
- type: FoVBackgroundModel
  datasets_names: [obs-1]
  spectral:
    type: PowerLawNormSpectralModel
    parameters:
    - {name: norm, value: 1.01, min: 0.0}
    - {name: tilt, value: 0.0, frozen: true}
    - {name: reference, value: 1.0, unit: TeV, frozen: true}
"""


def test_read_the_reference_layout():
    models = Models.from_yaml(REFERENCE_LAYOUT)
    assert models.names == ["source0", "source1", "obs-1-bkg"]
    s0, s1, bkg = models
    assert isinstance(s0.spatial_model, PointSpatialModel) and s0.spatial_model.frame == "galactic"
    assert isinstance(s0.spectral_model, ExpCutoffPowerLawSpectralModel)
    assert s0.spatial_model.lon_0.value == -0.5 and s0.spatial_model.lon_0.frozen and s0.spatial_model.lon_0.min == -180.0
    assert s0.spectral_model.lambda_.value == 0.006 and s0.spectral_model.lambda_.unit == "TeV-1"
    assert np.isnan(s0.spectral_model.index.min) and not s0.spectral_model.index.frozen
    assert s0.spectral_model.alpha.value == 1.0  # not in the file: the default, with a warning
    assert isinstance(s1.spatial_model, GaussianSpatialModel) and s1.spatial_model.sigma.min == 0.0
    assert s1.spectral_model.index.error == 0.1 and s1.spectral_model.reference.value == 1.0
    assert s1.datasets_names == ["obs-1"] and s1.apply_irf == {"exposure": True, "psf": False, "edisp": True}
    assert isinstance(bkg, FoVBackgroundModel) and bkg.datasets_names == ["obs-1"]
    assert bkg.spectral_model.norm.value == 1.01 and bkg.spectral_model.tilt.frozen
    for bad in ("TemplateNPredModel", "SkyDiffuseCube"):
        with pytest.raises(NotImplementedError, match=bad):
            Models.from_yaml(f"components:\n- name: x\n  type: {bad}\n")
    with pytest.raises(NotImplementedError, match="ConstantSpectralModel"):
        Models.from_yaml("components:\n- name: x\n  type: SkyModel\n  spectral: {type: ConstantSpectralModel, parameters: []}\n")


def _assert_same_models(a, b):
    assert a.names == b.names
    for m, n in zip(a, b):
        assert type(m) is type(n) and m.datasets_names == n.datasets_names
        assert type(m.spectral_model) is type(n.spectral_model)
        if getattr(m, "spatial_model", None) is not None:
            assert type(m.spatial_model) is type(n.spatial_model) and m.spatial_model.frame == n.spatial_model.frame
    for p, q in zip(a.parameters, b.parameters):
        assert (p.name, p.value, p.unit, p.frozen, p.error) == (q.name, q.value, q.unit, q.frozen, q.error)
        for x, y in ((p.min, q.min), (p.max, q.max)):
            assert (np.isnan(x) and np.isnan(y)) or x == y


@pytest.mark.parametrize("full_output", [False, True])
def test_round_trip_with_priors_links_and_limits(tmp_path, full_output):
    ds = configs.make_c3(width=(6.0, 3.0), n_sources=9)
    models = Models(ds.models)
    with_index = [m for m in models if "index" in m.spectral_model.parameters.names]
    with_index[0].spectral_model.index.prior = GaussianPrior(mu=2.2, sigma=0.3, weight=2)
    with_index[1].spectral_model.__dict__["index"] = with_index[0].spectral_model.index  # one parameter, two models
    models[1].spectral_model.amplitude.min = 0.0
    models[1].spectral_model.amplitude.prior = UniformPrior(min=0.0, max=1e-10)
    models[2].spatial_model.lon_0.frozen = True
    models[3].spectral_model.parameters["amplitude"].error = 3e-14
    models[4].apply_irf["psf"] = False
    path = tmp_path / "models.yaml"
    models.write(path, full_output=full_output)
    with pytest.raises(OSError):
        models.write(path)
    models.write(path, overwrite=True, full_output=full_output)
    back = Models.read(path)
    _assert_same_models(models, back)
    bi = [m for m in back if "index" in m.spectral_model.parameters.names]
    assert bi[1].spectral_model.index is bi[0].spectral_model.index
    assert len(back.parameters.unique_parameters) == len(models.parameters.unique_parameters)
    prior = bi[0].spectral_model.index.prior
    assert isinstance(prior, GaussianPrior) and (prior.mu, prior.sigma, prior.weight) == (2.2, 0.3, 2)
    uni = back[1].spectral_model.amplitude.prior
    assert isinstance(uni, UniformPrior) and (uni.min, uni.max) == (0.0, 1e-10)
    assert back[4].apply_irf == {"exposure": True, "psf": False, "edisp": True}
    text = path.read_text()
    assert ("scale_method" in text) == full_output and text.startswith("components:\n-   name: src-00\n    type: SkyModel\n")
    assert back.to_yaml(full_output) == models.to_yaml(full_output)  # a fixed point: same labels, same order


def test_models_read_back_give_the_same_npred():
    from oracle import adapter

    ds = configs.make_c3(width=(4.0, 2.0), n_sources=5, n_reco=6, n_true=12)
    want = adapter.evaluator_for(ds, conv="exact64").npred()
    ds.models = Models.from_yaml(Models(ds.models).to_yaml())
    got = adapter.evaluator_for(ds, conv="exact64").npred()
    np.testing.assert_array_equal(got, want)
    one = SkyModel.from_dict(ds.models[0].to_dict())
    assert one.name == ds.models[0].name and one.parameters.names == ds.models[0].parameters.names


def test_datasets_index_round_trip(tmp_path):
    """``Datasets.write / read`` (datasets/core.py:435-528): YAML index + one FITS file per dataset + the models YAML;
    every dataset gets back the models that name it (its background) and the ones that name nobody (the shared source)."""
    import yaml

    from gammapy_b200 import Datasets, MapDataset
    from oracle import adapter

    datasets = configs.make_c4(n_obs=3, width=(2.0, 2.0), n_reco=4, n_true=8)
    for d in datasets:
        d.counts = d.background  # something to compare
    datasets.write(tmp_path / "datasets.yaml", filename_models=tmp_path / "models.yaml")
    index = yaml.safe_load((tmp_path / "datasets.yaml").read_text())
    assert index == {"datasets": [{"name": d.name, "type": "MapDataset", "filename": d.name + ".fits"} for d in datasets]}
    with pytest.raises(OSError):
        datasets.write(tmp_path / "datasets.yaml")
    back = Datasets.read(tmp_path / "datasets.yaml", filename_models=tmp_path / "models.yaml")
    assert back.names == datasets.names and all(type(d) is MapDataset for d in back)
    assert back.models.names == datasets.models.names
    assert len(back.parameters) == len(datasets.parameters)
    shared = [m for m in back.models if m.datasets_names is None]
    assert len(shared) == 1 and all(d.models[shared[0].name] is shared[0] for d in back)
    for a, b in zip(datasets, back):
        assert [m.name for m in a.models] == [m.name for m in b.models]
        np.testing.assert_array_equal(a.exposure.data, b.exposure.data)
        np.testing.assert_array_equal(adapter.evaluator_for(a, conv="exact64").npred(),
                                      adapter.evaluator_for(b, conv="exact64").npred())
    (tmp_path / "bad.yaml").write_text("datasets:\n- {name: x, type: SpectrumDatasetOnOff, filename: x.fits}\n")
    with pytest.raises(NotImplementedError, match="SpectrumDatasetOnOff"):
        Datasets.read(tmp_path / "bad.yaml")


def test_on_off_dataset_fits_round_trip(tmp_path):
    from gammapy_b200 import Map, MapDatasetOnOff

    base = configs.make_dataset(width=(1.0, 1.0), n_reco=3, n_true=5, mask_radius=0.4, name="onoff")
    rng = np.random.RandomState(1)
    geom = base._geom
    ds = MapDatasetOnOff(counts=Map.from_geom(geom, data=rng.poisson(3.0, geom.data_shape).astype(float)),
                         counts_off=Map.from_geom(geom, data=rng.poisson(9.0, geom.data_shape).astype(float)),
                         acceptance=Map.from_geom(geom, data=np.ones(geom.data_shape)),
                         acceptance_off=Map.from_geom(geom, data=np.full(geom.data_shape, 3.0)),
                         exposure=base.exposure, psf=base.psf, edisp=base.edisp, mask_safe=base.mask_safe,
                         mask_fit=base.mask_fit, name="onoff")
    ds.write(tmp_path / "onoff.fits")
    back = MapDatasetOnOff.read(tmp_path / "onoff.fits")
    assert type(back) is MapDatasetOnOff and back.name == "onoff" and back.background is None
    for attr in ("counts", "counts_off", "acceptance", "acceptance_off", "exposure"):
        np.testing.assert_array_equal(getattr(ds, attr).data, getattr(back, attr).data)
    np.testing.assert_array_equal(back.alpha.data, np.full(geom.data_shape, 1.0 / 3.0))


def test_models_container_operations():
    """``Models`` as the reference's ``DatasetModels`` / ``Models`` container (modeling/models/core.py:776-790, 855-949,
    1046-1119, 1358-1380): unique names, selection, freezing by component, bounds, reassignment."""
    ds = configs.make_c3(width=(6.0, 3.0), n_sources=6)
    models = Models(ds.models)
    assert len(models) == 7 and models.names[-1] == "c3-bkg"
    points = models.select(tag="point", model_type="spatial")
    assert len(points) >= 1 and all(isinstance(m.spatial_model, PointSpatialModel) for m in points)
    assert models.select(tag="PointSpatialModel", model_type="spatial").names == points.names
    assert models.select(tag="fov-bkg").names == ["c3-bkg"] and models.select(name_substring="src-0").names == models.names[:6]
    assert models.select(datasets_names="other").names == models.names[:6]  # the background names its dataset
    assert models[models.selection_mask(tag="sky-model")].names == models.names[:6]
    with pytest.raises(ValueError, match="unique"):
        models.append(models[0])
    with pytest.raises(ValueError, match="unique"):
        models + models[1]
    extra = models[0].copy(name="extra")
    both = models + extra
    assert both.names[-1] == "extra" and len(models) == 7
    models.insert(0, extra)
    assert models.names[0] == "extra" and models.index("extra") == 0
    del models["extra"]
    assert "extra" not in models.names and len(models) == 7
    models.freeze("spatial")
    assert all(p.frozen for m in models[:6] for p in m.spatial_model.parameters)
    assert not models.frozen and not models[0].frozen
    assert models.select(frozen=True).names == []
    models.unfreeze("spatial")
    assert not models[0].spatial_model.lon_0.frozen
    models.freeze()
    assert models.frozen and models.select(frozen=True).names == models.names
    models.unfreeze()
    assert not models[0].spectral_model.amplitude.frozen and models[0].spectral_model.reference.frozen  # class defaults
    models.set_parameters_bounds(tag="pl", model_type="spectral", parameters_names="index", min=1.0, max=4.0)
    for m in models.select(tag="pl", model_type="spectral"):
        assert (m.spectral_model.index.min, m.spectral_model.index.max) == (1.0, 4.0)
    moved = models.reassign("c3", "c3-stacked")
    assert moved["c3-bkg"].datasets_names == ["c3-stacked"] and models["c3-bkg"].datasets_names == ["c3"]
    assert moved[0] is not models[0] and moved.names == models.names
    table = models.to_parameters_table()
    assert {r["model"] for r in table} == set(models.names) and table[0]["type"] == "spectral"
    text = str(models)
    assert text.startswith("Models\n\nComponent 0: SkyModel") and "lon_0" in text and "(frozen)" in text


def test_fit_covariance_on_the_models_and_in_the_side_file(tmp_path):
    """``Fit.run`` leaves the covariance where ``models.covariance`` finds it (modeling/fit.py:168-176, 286-292) and
    ``Models.write`` puts it into ``<stem>_covariance.dat`` in the reference's ``ascii.fixed_width`` layout
    (modeling/models/core.py:169-204, 716-757); read back, the matrix and the errors are the same.  The likelihood is a
    quadratic form evaluated on the CPU (``ShardedDatasets(local_stat=...)``), whose covariance is known exactly."""
    from gammapy_b200.distributed import ShardedDatasets
    from gammapy_b200.modeling import Fit

    model = SkyModel(spatial_model=PointSpatialModel(lon_0=0.1, lat_0=0.2, frame="galactic"),
                     spectral_model=ExpCutoffPowerLawSpectralModel(index=2.0, amplitude=1e-12, reference=1, lambda_=0.1),
                     name="src")
    model.spatial_model.lon_0.frozen = model.spatial_model.lat_0.frozen = True
    model.spectral_model.lambda_.frozen = True
    pars = list(model.parameters)
    names = [p.name for p in pars]
    hess = np.array([[80.0, 24.0], [24.0, 60.0]])  # second derivatives of the statistic in (index, ln amplitude)

    def local_stat(values):
        d = np.stack([values[:, names.index("index")] - 2.3, np.log(values[:, names.index("amplitude")] / 2e-12)], axis=1)
        return 0.5 * np.einsum("bi,ij,bj->b", d, hess, d)

    sharded = ShardedDatasets(None, local_stat=local_stat, global_models=[model], local_parameters=pars)
    result = Fit().run(sharded)
    assert result.success, result.message
    models = Models([model])
    cov = models.covariance
    assert cov.shape == (len(pars), len(pars)) and models.parameters_unique_names[0] == "src.spectral.index"
    i, a = names.index("index"), names.index("amplitude")
    amp = model.spectral_model.amplitude.value
    want = 2.0 * np.linalg.inv(hess) * np.outer([1.0, amp], [1.0, amp])  # errordef 1 on 2 log L; d amplitude = amp d ln
    # (the statistic is quadratic in ln(amplitude), Hesse differentiates in amplitude: agreement to the curvature of the log)
    np.testing.assert_allclose(cov[np.ix_([i, a], [i, a])], want, rtol=1e-2)
    np.testing.assert_allclose(np.sqrt(cov[i, i]), model.spectral_model.index.error, rtol=1e-12)
    frozen = [k for k, p in enumerate(pars) if p.frozen]
    assert np.all(cov[frozen][:, [i, a]] == 0) and np.all(cov[frozen, frozen] == 0)
    path = tmp_path / "fitted.yaml"
    models.write(path)
    text = (tmp_path / "fitted_covariance.dat").read_text().splitlines()
    assert text[0].replace(" ", "").startswith("|Parameters|0|1|") and text[1].replace(" ", "").startswith("|src.spectral.index|")
    assert "covariance: fitted_covariance.dat" in path.read_text()
    back = Models.read(path)
    np.testing.assert_array_equal(back.covariance, cov)
    assert back[0].spectral_model.index.error == model.spectral_model.index.error
    # a copy keeps the errors, not the correlations; files without covariance are read as before
    clone = models.copy()
    assert clone.covariance[i, a] == 0 and clone.covariance[i, i] == pytest.approx(cov[i, i])
    models.write(tmp_path / "plain.yaml", write_covariance=False)
    assert "covariance" not in (tmp_path / "plain.yaml").read_text()
    (tmp_path / "fitted_covariance.dat").write_text("garbage\n")
    assert Models.read(path).covariance[i, a] == 0  # unreadable side file: a warning, diagonal from the errors
