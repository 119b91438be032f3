"""CAR geometries about a reference point off the equator -- what ``WcsGeom.create(skydir=(lon, lat))`` builds
(maps/wcs/geom.py:370-410: CRVAL = skydir), an oblique plate carree in wcslib.  The product tilts the native origin up
to the reference latitude (one rotation about the y axis); the oracle restates FITS WCS paper II in its general form
(native pole from eqs. 8-10 with the default LONPOLE / LATPOLE, rotations eqs. 2 and 5).  The two formulations must
agree, and both must have the properties of the projection.  wcslib / astropy cannot be run here: parity with the
reference's wcs_pix2world is by these formulas, unpinned."""
import numpy as np
import pytest

from gammapy_b200.maps import WcsGeom
from gammapy_b200.utils import DEG, angular_separation
from oracle.wcs import ImageGeom


@pytest.mark.parametrize("skydir", [(83.63, 22.01), (266.4, -29.0), (10.0, 1e-3), (359.9, 75.0), (0.5, -60.0)])
def test_oblique_car_product_vs_oracle_and_analytic(skydir):
    g = WcsGeom.create(skydir=skydir, binsz=0.05, width=(6.0, 4.0), proj="CAR", frame="icrs")
    o = ImageGeom.create(skydir=skydir, binsz=0.05, npix=(120, 80), proj="CAR")
    assert g.npix == (120, 80) and g.crval == skydir and g.crpix == (60.5, 40.5)
    c = g.get_coord()
    lon, lat = o.get_coord()
    dlon = (c["lon"] - lon + 180.0) % 360.0 - 180.0
    assert np.abs(dlon).max() < 1e-11 and np.abs(c["lat"] - lat).max() < 1e-12
    yy, xx = np.mgrid[0:80, 0:120]
    px, py = g.coord_to_pix((c["lon"], c["lat"]))
    assert np.abs(px - xx).max() < 1e-9 and np.abs(py - yy).max() < 1e-9
    opx, opy = o.world_to_pix(lon, lat)
    assert np.abs(opx - xx).max() < 1e-8 and np.abs(opy - yy).max() < 1e-8
    # the rotation keeps distances: a pixel is as far from the reference point as its native coordinates are from (0, 0)
    sep = angular_separation(c["lon"] * DEG, c["lat"] * DEG, skydir[0] * DEG, skydir[1] * DEG)
    native = angular_separation(-(xx - 59.5) * 0.05 * DEG, (yy - 39.5) * 0.05 * DEG, 0.0, 0.0)
    np.testing.assert_allclose(sep, native, atol=1e-14)
    # ... and angles at the reference point: north is up, east is left, the reference meridian is the central column
    assert c["lat"][-1, 60] > c["lat"][0, 60]
    assert ((c["lon"][40, 0] - c["lon"][40, -1] + 180.0) % 360.0 - 180.0) > 0
    mid = 0.5 * (c["lon"][:, 59] + c["lon"][:, 60])
    assert np.abs((mid - skydir[0] + 180.0) % 360.0 - 180.0).max() < 1e-9
    # the centre of the image is the reference point, so kernels and cutouts built about it (to_odd_npix) keep it
    assert abs(g.center_skydir.lon - skydir[0]) < 1e-12 and abs(g.center_skydir.lat - skydir[1]) < 1e-12
    # solid angle: same planar-triangle formula on the corners (geom.py:803-854); the native pixel at latitude theta
    # covers binsz^2 cos(theta) wherever the origin was rotated to
    omega = g.solid_angle()
    np.testing.assert_allclose(omega, o.solid_angle(), rtol=1e-9)
    theta = (yy - 39.5) * 0.05 * DEG
    np.testing.assert_allclose(omega, (0.05 * DEG) ** 2 * np.cos(theta), rtol=2e-3)
    # cutouts follow the same integer rules
    ys, xs = g.cutout_slices_for((c["lon"][30, 50], c["lat"][30, 50]), 1.0, odd_npix=True)
    _, (oys, oxs) = o.cutout_info((lon[30, 50], lat[30, 50]), 1.0, odd_npix=True)
    assert (ys, xs) == (oys, oxs) and xs.stop - xs.start == 21 and ys == slice(20, 41)


def test_oblique_car_tends_to_the_equatorial_map():
    """crval latitude -> 0: the separable linear map, continuously; and the native pole the oracle computes."""
    eq = WcsGeom.create(skydir=(30.0, 0.0), binsz=0.02, width=(2.0, 2.0), proj="CAR", frame="galactic")
    ob = WcsGeom.create(skydir=(30.0, 1e-9), binsz=0.02, width=(2.0, 2.0), proj="CAR", frame="galactic")
    a, b = eq.get_coord(), ob.get_coord()
    assert np.abs(a["lon"] - b["lon"]).max() < 1e-9 and np.abs(a["lat"] - b["lat"]).max() < 2e-9
    for lat0, want in ((22.01, (83.63 + 180.0, 90.0 - 22.01, 0.0)), (-29.0, (83.63, 90.0 - 29.0, 180.0))):
        alpha_p, delta_p, phi_p = ImageGeom.create(skydir=(83.63, lat0), binsz=0.1, npix=(10, 10))._car_pole()
        got = ((alpha_p / DEG) % 360.0, delta_p / DEG, phi_p / DEG)
        np.testing.assert_allclose(got, (want[0] % 360.0, want[1], want[2]), atol=1e-10)
    with pytest.raises(ValueError):
        WcsGeom.create(skydir=(0.0, 90.0), binsz=0.1, width=1.0, proj="CAR")


def test_a_shifted_reference_pixel_is_not_a_shifted_reference_point():
    """An equatorial CAR image whose reference pixel lies far below the image (what a cutout of a tall map is) and an
    oblique CAR about the same centre are different pixelisations of the sky: the first has pixels compressed in
    longitude by cos(lat), the second has square pixels at its centre."""
    shifted = WcsGeom((100, 100), 0.02, (30.0, 0.0), (50.5, 50.5 - 20.0 / 0.02), frame="galactic")
    oblique = WcsGeom.create(skydir=(30.0, 20.0), binsz=0.02, width=(2.0, 2.0), proj="CAR", frame="galactic")
    assert abs(shifted.center_skydir.lat - 20.0) < 1e-12 and abs(oblique.center_skydir.lat - 20.0) < 1e-12
    ratio = shifted.solid_angle()[50, 50] / oblique.solid_angle()[50, 50]
    assert abs(ratio - np.cos(20.0 * DEG)) < 1e-4


def test_reference_spatial_model_checks_on_oblique_geometries():
    """The reference's data-free spatial-model checks that live on geometries off the equator
    (modeling/models/tests/test_spatial.py:40-59 ``test_sky_point_source``: a point source at (2.5, 2.5) on a map about
    (2.4, 2.3) integrates to 1; :75-88 ``test_sky_gaussian``: an elongated Gaussian about (2, 2) is normalised to 1e-3
    over a 20 deg map -- there on an AIT map, here on the oblique CAR map the same ``WcsGeom.create`` call builds by
    default): coordinates and solid angles of the oblique projection, host mirror and oracle."""
    from gammapy_b200.modeling.models import GaussianSpatialModel
    from oracle import spatial

    og = ImageGeom.create(skydir=(2.4, 2.3), npix=(10, 10), binsz=0.3)
    w = spatial.point_integrate_geom({"type": "point", "lon_0": 2.5, "lat_0": 2.5}, og)
    assert_sum = w.sum()
    assert abs(assert_sum - 1.0) < 1e-12 and np.count_nonzero(w) == 4
    px, py = og.world_to_pix(2.5, 2.5)
    iy, ix = np.unravel_index(np.argmax(w), w.shape)
    assert abs(ix - px) <= 0.5 and abs(iy - py) <= 0.5
    g = WcsGeom.create(skydir=(2.4, 2.3), npix=(10, 10), binsz=0.3, frame="icrs")
    gx, gy = g.coord_to_pix((2.5, 2.5))
    assert abs(gx - px) < 1e-10 and abs(gy - py) < 1e-10

    big = WcsGeom.create(binsz=0.05, width=(20, 20), skydir=(2, 2), frame="galactic")
    obig = ImageGeom.create(skydir=(2.0, 2.0), binsz=0.05, npix=(400, 400))
    model = GaussianSpatialModel(lon_0=2, lat_0=2, sigma=3, e=0.8, phi=30, frame="galactic")
    c = big.get_coord()
    total = np.sum(model(c["lon"], c["lat"]) * big.solid_angle())
    assert abs(total - 1.0) < 1e-3
    lon, lat = obig.get_coord()
    ototal = np.sum(spatial.gauss_evaluate(lon, lat, 2.0, 2.0, 3.0, 0.8, 30.0) * obig.solid_angle())
    assert abs(ototal - 1.0) < 1e-3 and abs(total - ototal) < 1e-9
    assert model.evaluation_radius == 15.0
