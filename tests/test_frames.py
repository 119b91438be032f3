"""ICRS <-> Galactic: the product's rotation matrix (gammapy_b200/utils.py, the same numbers in csrc/gpy_b200.cu) against
the oracle's spherical-trigonometry restatement (oracle/frames.py) and against the two positions astropy is known to
print; models in another frame than the map through the oracle.  astropy cannot be imported here: parity with its
frame graph is unpinned beyond those positions."""
import numpy as np
import pytest

from gammapy_b200 import configs
from gammapy_b200.maps import WcsGeom
from gammapy_b200.utils import ICRS_TO_GALACTIC, SkyCoord, transform_lonlat
from oracle import frames


def test_known_positions_and_two_formulations():
    # SkyCoord(0, 0, frame="galactic").icrs and SkyCoord(0, 90, frame="galactic").icrs as astropy prints them
    for fn in (lambda l, b: transform_lonlat(l, b, "galactic", "icrs"), frames.galactic_to_icrs):
        np.testing.assert_allclose(fn(0.0, 0.0), (266.40498829, -28.93617776), atol=5e-9)
        np.testing.assert_allclose(fn(0.0, 90.0), (192.85947789, 27.12825241), atol=5e-9)
    # the Crab: ICRS (83.63308, 22.0145) is (184.55745, -5.78436) in galactic coordinates
    np.testing.assert_allclose(transform_lonlat(83.63308, 22.0145, "icrs", "galactic"), (184.55745, -5.78436), atol=2e-5)
    np.testing.assert_allclose(ICRS_TO_GALACTIC @ ICRS_TO_GALACTIC.T, np.eye(3), atol=1e-15)
    rng = np.random.RandomState(0)
    lon, lat = rng.uniform(0, 360, 2000), np.degrees(np.arcsin(rng.uniform(-1, 1, 2000)))
    for src, dst, fn in (("icrs", "galactic", frames.icrs_to_galactic), ("galactic", "icrs", frames.galactic_to_icrs)):
        a, b = transform_lonlat(lon, lat, src, dst), fn(lon, lat)
        scale = np.cos(np.radians(b[1]))
        assert np.abs(((a[0] - b[0] + 180.0) % 360.0 - 180.0) * scale).max() < 1e-12 and np.abs(a[1] - b[1]).max() < 1e-12
    back = transform_lonlat(*transform_lonlat(lon, lat, "icrs", "galactic"), "galactic", "icrs")
    assert np.abs(((back[0] - lon + 180.0) % 360.0 - 180.0) * np.cos(np.radians(lat))).max() < 1e-12
    with pytest.raises(NotImplementedError):
        transform_lonlat(0.0, 0.0, "icrs", "fk4")


def test_positions_enter_a_geometry_in_its_own_frame():
    geom = WcsGeom.create(skydir=(184.55, -5.78), binsz=0.02, width=(2.0, 2.0), frame="galactic")
    crab = SkyCoord(83.63308, 22.0145, frame="icrs")
    px, py = geom.coord_to_pix(crab)
    want = geom.coord_to_pix(transform_lonlat(83.63308, 22.0145, "icrs", "galactic"))
    assert (px, py) == (want[0], want[1]) and abs(px - 49.5) < 1.0 and abs(py - 49.5) < 1.0
    assert bool(geom.contains(crab)) and not bool(geom.contains(SkyCoord(83.63308, 22.0145, frame="galactic")))
    assert geom.separation(crab)[50, 50] < 0.03
    assert geom.cutout_slices_for(crab, 0.5) == geom.cutout_slices_for(crab.galactic, 0.5)
    assert crab.galactic.icrs.separation(crab) < 1e-12 and crab.separation(crab.galactic) < 1e-12


def _dataset(model_frame, skydir, proj, map_frame="galactic"):
    from gammapy_b200.modeling.models import (DiskSpatialModel, FoVBackgroundModel, GaussianSpatialModel,
                                              PointSpatialModel, PowerLawSpectralModel, SkyModel)

    ds = configs.make_dataset(width=(3.0, 2.0), n_reco=5, n_true=10, skydir=skydir, proj=proj, mask_radius=None, name="f")
    assert ds._geom.frame == map_frame
    cosb = np.cos(np.radians(skydir[1]))
    # (offsets that keep every position away from pixel edges: there the cutout's rounding hangs on the last bit)
    pos = [(skydir[0] + 0.6043 / cosb, skydir[1] + 0.3131), (skydir[0] - 0.5017 / cosb, skydir[1] - 0.2071),
           (skydir[0] + 0.1033 / cosb, skydir[1] - 0.4029)]
    if model_frame != map_frame:
        pos = [tuple(float(v) for v in transform_lonlat(l, b, map_frame, model_frame)) for l, b in pos]

    def pl():
        return PowerLawSpectralModel(index=2.2, amplitude="3e-12 cm-2 s-1 TeV-1", reference="1 TeV")

    ds.models = [
        SkyModel(spatial_model=PointSpatialModel(lon_0=pos[0][0], lat_0=pos[0][1], frame=model_frame), spectral_model=pl(), name="p"),
        SkyModel(spatial_model=GaussianSpatialModel(lon_0=pos[1][0], lat_0=pos[1][1], sigma=0.15, frame=model_frame),
                 spectral_model=pl(), name="g"),
        SkyModel(spatial_model=DiskSpatialModel(lon_0=pos[2][0], lat_0=pos[2][1], r_0=0.25, frame=model_frame),
                 spectral_model=pl(), name="d"),
        FoVBackgroundModel(dataset_name="f")]
    return ds


@pytest.mark.parametrize("skydir,proj", [((0.0, 0.0), "CAR"), ((184.55, -5.78), "CAR"), ((120.0, 35.0), "TAN")])
def test_oracle_round_models_do_not_care_which_frame_names_them(skydir, proj):
    """Circular sources at the same sky positions, written in galactic or in ICRS coordinates, on a galactic map: same
    cutouts and the same npred up to the rounding of the coordinate rotation."""
    from oracle import adapter

    a = adapter.evaluator_for(_dataset("galactic", skydir, proj), conv="exact64")
    b = adapter.evaluator_for(_dataset("icrs", skydir, proj), conv="exact64")
    na, nb = a.npred(), b.npred()
    assert np.max(np.abs(na - nb) / np.maximum(na, 1e-300)) < 1e-9
    for name in ("p", "g", "d"):
        assert a.evaluators[name].slices == b.evaluators[name].slices
