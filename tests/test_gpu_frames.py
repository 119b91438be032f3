"""Spatial models in another celestial frame than the map (ICRS models on a galactic map and the reverse): the device
rotates pixel coordinates into the model frame (``to_model_frame`` in kernels.cuh), the host brings the position into
the map frame for cutouts and IRF lookups (``frame_convert`` in gpy_b200.cu); GPU vs the oracle, whose transform is
written from the same published definition in its other (spherical trigonometry) form (oracle/frames.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sources(frame, positions, elongated=True):
    from gammapy_b200.modeling.models import (DiskSpatialModel, ExpCutoffPowerLawSpectralModel, GaussianSpatialModel,
                                              LogParabolaSpectralModel, PointSpatialModel, PowerLawSpectralModel, SkyModel)

    e, phi = (0.7, 35.0) if elongated else (0.0, 0.0)
    return [
        SkyModel(spatial_model=PointSpatialModel(lon_0=positions[0][0], lat_0=positions[0][1], frame=frame),
                 spectral_model=PowerLawSpectralModel(index=2.2, amplitude=3e-11, reference=1), name="point"),
        SkyModel(spatial_model=GaussianSpatialModel(lon_0=positions[1][0], lat_0=positions[1][1], sigma=0.2, e=e, phi=phi,
                                                    frame=frame),
                 spectral_model=LogParabolaSpectralModel(amplitude=4e-11, reference=1, alpha=2.1, beta=0.2), name="gauss"),
        SkyModel(spatial_model=DiskSpatialModel(lon_0=positions[2][0], lat_0=positions[2][1], r_0=0.3, e=e, phi=phi + 40.0,
                                                frame=frame),
                 spectral_model=ExpCutoffPowerLawSpectralModel(index=1.9, amplitude=5e-11, reference=1, lambda_=0.1),
                 name="disk"),
    ]


@pytest.mark.parametrize("skydir,proj", [((0.0, 0.0), "CAR"), ((184.55, -5.78), "CAR"), ((120.0, 35.0), "TAN")])
def test_icrs_models_on_a_galactic_map(ctx, skydir, proj):
    from gammapy_b200 import Map, configs
    from gammapy_b200.modeling.models import FoVBackgroundModel
    from gammapy_b200.utils import transform_lonlat
    from oracle import adapter

    ds = configs.make_dataset(width=(4.0, 3.0), skydir=skydir, proj=proj, mask_radius=1.3, name="frames")
    cosb = np.cos(np.radians(skydir[1]))
    gal = [(skydir[0] + 0.7043 / cosb, skydir[1] + 0.4131), (skydir[0] - 0.6017 / cosb, skydir[1] - 0.3071),
           (skydir[0] + 0.1033 / cosb, skydir[1] - 0.6029)]
    icrs = [tuple(float(v) for v in transform_lonlat(l, b, "galactic", "icrs")) for l, b in gal]
    ds.models = _sources("icrs", icrs) + [FoVBackgroundModel(dataset_name="frames")]
    de = adapter.evaluator_for(ds, conv="exact64")
    got, want = ds.npred().data, de.npred()
    assert np.max(np.abs(got - want) / np.maximum(want, 1e-300)) < 1e-5
    for m in ds.models[:3]:
        st = ds.evaluator_state(m.name)
        ys, xs = de.evaluators[m.name].slices
        assert (st["x0"], st["y0"], st["cx"], st["cy"]) == (xs.start, ys.start, xs.stop - xs.start, ys.stop - ys.start)
    ds.counts = Map.from_geom(ds._geom, data=de.fake(6), dtype=float)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    # circular sources: the frame that names the position does not matter
    ds.models = _sources("icrs", icrs, elongated=False) + [FoVBackgroundModel(dataset_name="frames")]
    in_icrs = ds.npred().data
    ds.models = _sources("galactic", gal, elongated=False) + [FoVBackgroundModel(dataset_name="frames")]
    in_gal = ds.npred().data
    assert np.max(np.abs(in_icrs - in_gal) / np.maximum(in_gal, 1e-300)) < 1e-6
    # ... elongated ones keep their position angle in their own frame: rotating the frame rotates the ellipse
    ds.models = _sources("galactic", gal) + [FoVBackgroundModel(dataset_name="frames")]
    assert np.max(np.abs(ds.npred().data - got) / np.maximum(got, 1e-300)) > 1e-2
    # a source moves in ICRS coordinates: new cutout, same agreement
    ds.models = _sources("icrs", icrs) + [FoVBackgroundModel(dataset_name="frames")]
    de = adapter.evaluator_for(ds, conv="exact64")
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    ds.models["gauss"].spatial_model.lon_0.value += 0.3
    ds.models["gauss"].spatial_model.lat_0.value -= 0.2
    adapter.refresh_models(de, ds)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3


def test_galactic_models_on_an_icrs_map_with_irf_maps(ctx, monkeypatch):
    """The reverse direction, on an oblique CAR map in ICRS (a Crab field) with position-dependent IRFs: the lookups
    happen at the position in the map's frame."""
    from gammapy_b200 import Map, configs
    from gammapy_b200.modeling.models import FoVBackgroundModel
    from gammapy_b200.utils import transform_lonlat
    from oracle import adapter

    monkeypatch.setattr(configs, "FRAME", "icrs")
    skydir = (83.63, 22.01)
    ds = configs.make_dataset(width=(4.0, 2.0), n_reco=6, n_true=12, skydir=skydir, mask_radius=None, name="crab")
    assert ds._geom.frame == "icrs"
    configs.zoned_irfs(ds)
    cosb = np.cos(np.radians(skydir[1]))
    icrs = [(skydir[0] + 1.2043 / cosb, skydir[1] + 0.3131), (skydir[0] - 1.1017 / cosb, skydir[1] - 0.2071),
            (skydir[0] + 0.2033 / cosb, skydir[1] - 0.3029)]
    gal = [tuple(float(v) for v in transform_lonlat(l, b, "icrs", "galactic")) for l, b in icrs]
    ds.models = _sources("galactic", gal) + [FoVBackgroundModel(dataset_name="crab")]
    de = adapter.evaluator_for(ds, conv="exact64")
    got, want = ds.npred().data, de.npred()
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 1e-5
    assert ds.evaluator_irf("point", False)["edisp_pixel"] != ds.evaluator_irf("gauss", False)["edisp_pixel"]
    ds.counts = Map.from_geom(ds._geom, data=de.fake(2), dtype=float)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3


def test_integrate_geom_and_unknown_frames(ctx):
    from gammapy_b200 import WcsGeom
    from gammapy_b200.modeling.models import GaussianSpatialModel
    from gammapy_b200.utils import transform_lonlat

    geom = WcsGeom.create(skydir=(184.55, -5.78), binsz=0.02, width=(2.0, 2.0), frame="galactic")
    lon, lat = (float(v) for v in transform_lonlat(184.6043, -5.7371, "galactic", "icrs"))
    a = GaussianSpatialModel(lon_0=184.6043, lat_0=-5.7371, sigma=0.1, frame="galactic").integrate_geom(geom).data
    b = GaussianSpatialModel(lon_0=lon, lat_0=lat, sigma=0.1, frame="icrs").integrate_geom(geom).data
    np.testing.assert_allclose(b, a, rtol=1e-6, atol=1e-9 * a.max())
    with pytest.raises(NotImplementedError, match="fk5"):
        GaussianSpatialModel(lon_0=lon, lat_0=lat, sigma=0.1, frame="fk5").integrate_geom(geom)
