"""IRF kernel extraction at the source position (PSFMap.get_psf_kernel irf/psf/map.py:249-337,
EDispKernelMap.get_edisp_kernel irf/edisp/map.py:369-401) with the refresh-on-drift logic of MapEvaluator.update
(datasets/evaluator.py:174-243, 465-481), GPU vs the oracle restatement (oracle/irf.py; parity unpinned beyond the
position-independent limit, which tests/test_oracle_irf.py ties to the kernel the reference's own npred golden pins)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _zoned(models, **kw):
    from gammapy_b200 import configs

    ds = configs.make_dataset(width=(4.0, 2.0), n_reco=6, n_true=12, mask_radius=None, **kw)
    configs.zoned_irfs(ds)
    ds.models = models
    return ds


def _point(lon, lat, name="src", amplitude=2e-11):
    from gammapy_b200 import configs
    from gammapy_b200.modeling.models import PointSpatialModel, PowerLawSpectralModel, SkyModel

    return SkyModel(spatial_model=PointSpatialModel(lon_0=lon, lat_0=lat, frame=configs.FRAME),
                    spectral_model=PowerLawSpectralModel(index=2.2, amplitude=amplitude, reference=1), name=name)


def _check(ds, de, rtol=1e-5):
    got, want = ds.npred().data, de.npred()
    rel = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
    assert rel < rtol, rel
    return got


def test_kernel_follows_the_source_across_a_zone_boundary(ctx):
    from gammapy_b200.modeling.models import FoVBackgroundModel
    from oracle import adapter, irf as oirf

    src = _point(1.0, 0.1)
    ds = _zoned([src, FoVBackgroundModel(dataset_name="c1")])
    de = adapter.evaluator_for(ds, conv="exact64")
    od = de.ds
    a = _check(ds, de)
    st = ds.evaluator_irf("src")
    want = oirf.psf_kernel_at(od.psf_map, od.rad_edges, od.irf_geom, (1.0, 0.1), od.geom.binsz, tuple(od.geom.center_skydir))
    assert st["n_updates"] == 1 and st["k"] == want.shape[-1]
    np.testing.assert_allclose(st["psf_kernel"], want, rtol=0, atol=2e-7 * want.max())
    np.testing.assert_allclose(st["psf_kernel"].sum(axis=(1, 2)), 1.0, rtol=1e-5)
    pix_a = st["edisp_pixel"]

    # The reference's evaluator starts with _cached_position = (0, 0) (evaluator.py:93) and its first update does not
    # touch it, so the first spatial step looks the IRFs up once more, however small it is; from then on a step inside
    # evaluation_radius + 0.1 deg of the last lookup re-uses them.  Same on both sides.
    src.spatial_model.lon_0.value = 0.95
    adapter.refresh_models(de, ds)
    _check(ds, de)
    assert ds.evaluator_irf("src", with_kernel=False)["n_updates"] == 2
    assert de.evaluators["src"].n_irf_updates == 2
    src.spatial_model.lon_0.value = 0.99
    src.spatial_model.lat_0.value = 0.12
    adapter.refresh_models(de, ds)
    _check(ds, de)
    assert ds.evaluator_irf("src", with_kernel=False)["n_updates"] == 2
    assert de.evaluators["src"].n_irf_updates == 2
    src.spatial_model.lat_0.value = 0.1

    # across the meridian into the other zone: wider PSF, another edisp matrix
    src.spatial_model.lon_0.value = -1.0
    adapter.refresh_models(de, ds)
    b = _check(ds, de)
    st2 = ds.evaluator_irf("src")
    want2 = oirf.psf_kernel_at(od.psf_map, od.rad_edges, od.irf_geom, (-1.0, 0.1), od.geom.binsz, tuple(od.geom.center_skydir))
    assert st2["n_updates"] == 3 and de.evaluators["src"].n_irf_updates == 3
    assert st2["k"] == want2.shape[-1] and st2["k"] > st["k"]
    np.testing.assert_allclose(st2["psf_kernel"], want2, rtol=0, atol=2e-7 * want2.max())
    assert st2["edisp_pixel"] != pix_a
    assert abs(a.sum() - b.sum()) > 10.0  # the zones really differ (the edisp bias moves counts across the axis edge)

    # in the transition column the profile is the bilinear mix of the two zones
    src.spatial_model.lon_0.value = 0.2
    src.spatial_model.lat_0.value = -0.3
    adapter.refresh_models(de, ds)
    _check(ds, de)
    st3 = ds.evaluator_irf("src")
    want3 = oirf.psf_kernel_at(od.psf_map, od.rad_edges, od.irf_geom, (0.2, -0.3), od.geom.binsz, tuple(od.geom.center_skydir))
    assert st3["k"] == want3.shape[-1]
    np.testing.assert_allclose(st3["psf_kernel"], want3, rtol=0, atol=2e-7 * want3.max())


def test_two_sources_in_two_zones_and_the_statistic(ctx):
    """Two evaluators with their own PSF kernel and edisp matrix in one dataset (two edisp groups in the fold kernel),
    an extended source among them; npred per bin and the Cash statistic against the oracle."""
    from gammapy_b200 import Map, configs
    from gammapy_b200.modeling.models import FoVBackgroundModel, GaussianSpatialModel, PowerLawSpectralModel, SkyModel
    from oracle import adapter

    g = SkyModel(spatial_model=GaussianSpatialModel(lon_0=-1.1, lat_0=-0.2, sigma=0.15, frame=configs.FRAME),
                 spectral_model=PowerLawSpectralModel(index=2.6, amplitude=3e-11, reference=1), name="gauss")
    ds = _zoned([_point(1.2, 0.3, "point"), g, FoVBackgroundModel(dataset_name="c1")])
    de = adapter.evaluator_for(ds, conv="exact64")
    _check(ds, de)
    assert ds.evaluator_irf("point", False)["edisp_pixel"] != ds.evaluator_irf("gauss", False)["edisp_pixel"]
    assert ds.evaluator_irf("point", False)["k"] < ds.evaluator_irf("gauss", False)["k"]
    counts = de.fake(5)
    ds.counts = Map.from_geom(ds._geom, data=counts, dtype=float)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    # steady state: a spectral step re-uses the kernels and the convolved cubes
    ds.models["point"].spectral_model.index.value = 2.4
    adapter.refresh_models(de, ds)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    assert ds.evaluator_irf("point", False)["n_updates"] == 1


def test_batched_call_refuses_a_lookup(ctx):
    from gammapy_b200 import Map
    from gammapy_b200.modeling.models import FoVBackgroundModel

    src = _point(1.0, 0.0)
    ds = _zoned([src, FoVBackgroundModel(dataset_name="c1")])
    ds.counts = Map.from_geom(ds._geom, data=np.zeros(ds._geom.data_shape), dtype=float)
    eng = ds._get_engine()
    base = eng.values()
    ds.stat_sum()
    col = [i for i, p in enumerate(eng.parameters) if p.name == "lon_0"][0]
    grid = np.tile(base, (3, 1))
    grid[1, col] += 0.01   # inside the support of the last lookup: the vector uses the evaluator's current IRFs
    pair = eng.stat_sum(grid[:2])
    assert np.all(np.isfinite(pair)) and pair[0] == ds.stat_sum()
    grid[2, col] -= 2.0    # would need MapEvaluator.update with a new IRF lookup
    with pytest.raises(NotImplementedError):
        eng.stat_sum(grid)


@pytest.mark.parametrize("proj,skydir", [("CAR", (83.63, 22.01)), ("TAN", (266.4, -29.0))])
def test_irf_lookup_off_the_equator(ctx, proj, skydir):
    """IRF maps and dataset on an oblique CAR / TAN geometry (what MapDatasetMaker produces for a field off the
    equator): the lookup pixel follows the IRF grid's projection, the kernel geometry is created about the dataset's
    centre in the dataset's projection (geom.to_odd_npix, maps/wcs/geom.py:1106-1138), so its pixel separations are
    those of the native coordinates -- not those of an equatorial grid at that latitude."""
    from gammapy_b200 import Map, configs
    from gammapy_b200.modeling.models import FoVBackgroundModel, GaussianSpatialModel, PowerLawSpectralModel, SkyModel
    from oracle import adapter, irf as oirf

    cosb = np.cos(np.radians(skydir[1]))
    p_pos = (skydir[0] + 1.2 / cosb, skydir[1] + 0.3)
    g_pos = (skydir[0] - 1.1 / cosb, skydir[1] - 0.2)
    g = SkyModel(spatial_model=GaussianSpatialModel(lon_0=g_pos[0], lat_0=g_pos[1], sigma=0.15, frame=configs.FRAME),
                 spectral_model=PowerLawSpectralModel(index=2.6, amplitude=3e-11, reference=1), name="gauss")
    ds = _zoned([_point(p_pos[0], p_pos[1], "point"), g, FoVBackgroundModel(dataset_name="c1")], skydir=skydir, proj=proj)
    assert ds.psf.geom.projection == proj and ds.psf.geom.crval[1] == pytest.approx(skydir[1])
    de = adapter.evaluator_for(ds, conv="exact64")
    od = de.ds
    _check(ds, de)
    for name, pos in (("point", p_pos), ("gauss", g_pos)):
        st = ds.evaluator_irf(name)
        want = oirf.psf_kernel_at(od.psf_map, od.rad_edges, od.irf_geom, pos, od.geom.binsz, tuple(od.geom.center_skydir),
                                  proj=proj)
        assert st["k"] == want.shape[-1] and st["n_updates"] == 1
        np.testing.assert_allclose(st["psf_kernel"], want, rtol=0, atol=2e-7 * want.max())
    assert ds.evaluator_irf("point", False)["edisp_pixel"] != ds.evaluator_irf("gauss", False)["edisp_pixel"]
    assert ds.evaluator_irf("point", False)["k"] < ds.evaluator_irf("gauss", False)["k"]
    ds.counts = Map.from_geom(ds._geom, data=de.fake(5), dtype=float)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    # the point source crosses into the other zone
    ds.models["point"].spatial_model.lon_0.value = skydir[0] - 0.9 / cosb
    adapter.refresh_models(de, ds)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    assert ds.evaluator_irf("point", False)["edisp_pixel"] == ds.evaluator_irf("gauss", False)["edisp_pixel"] or \
        ds.evaluator_irf("point", False)["k"] == ds.evaluator_irf("gauss", False)["k"]


@pytest.mark.parametrize("n_reco,n_true", [(6, 12), (30, 60)])
def test_every_source_with_its_own_edisp_matrix(ctx, n_reco, n_true):
    """IRF maps that vary from IRF pixel to IRF pixel (graded_irfs: what MapDatasetMaker fills for a real observation):
    nine sources, nine PSF kernels, as many edisp matrices as the sources occupy IRF pixels -- more than any staged fold
    kernel holds, so the any-shape kernel runs and works through the matrices of each tile's own sources one at a time.
    npred per bin, the statistic, a moved source and a batched scan against the oracle."""
    from gammapy_b200 import Map, configs
    from oracle import adapter

    ds = configs.make_c3(width=(6.0, 3.0), n_sources=9, n_reco=n_reco, n_true=n_true)
    configs.graded_irfs(ds)
    de = adapter.evaluator_for(ds, conv="exact64")
    _check(ds, de)
    names = [m.name for m in ds.models if getattr(m, "spatial_model", None) is not None]
    pixels = {ds.evaluator_irf(n, False)["edisp_pixel"] for n in names}
    assert len(pixels) >= 6, pixels
    ds.counts = Map.from_geom(ds._geom, data=de.fake(3), dtype=float)
    want = de.stat_sum()
    assert abs(ds.stat_sum() - want) < 1e-3
    # a spectral step (steady state) and a source that moves into another IRF pixel
    ds.models[names[0]].spectral_model.parameters["amplitude"].value *= 1.3
    adapter.refresh_models(de, ds)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    ds.models[names[1]].spatial_model.lon_0.value += 0.6
    adapter.refresh_models(de, ds)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    for n in names:  # the same evaluators looked their IRFs up again on both sides (the drift rule of MapEvaluator.update)
        assert ds.evaluator_irf(n, False)["n_updates"] == de.evaluators[n].n_irf_updates, n
    # batched scan from this state: vector 0 is the single evaluation, the others move a spectral parameter
    eng = ds._get_engine()
    grid = np.tile(eng.values(), (3, 1))
    col = [i for i, p in enumerate(eng.parameters) if p.name in ("index", "alpha")][0]
    grid[:, col] += [0.0, 0.05, 0.1]
    scan = eng.stat_sum(grid)
    assert abs(scan[0] - de.stat_sum()) < 1e-3 and scan[1] != scan[0] and scan[2] != scan[1]
