"""TAN (gnomonic) image geometries: the product's WcsGeom (direct gnomonic formulas) against the oracle's ImageGeom
(native spherical coordinates + rotation, FITS WCS paper II eqs. 2, 5, 14, 15, 54, 55) and analytic properties.
wcslib / astropy cannot be run here: parity with the reference's wcs_pix2world is by these formulas, unpinned."""
import numpy as np
import pytest

from gammapy_b200.maps import WcsGeom
from gammapy_b200.utils import DEG, angular_separation
from oracle.wcs import ImageGeom


@pytest.mark.parametrize("skydir", [(83.63, 22.01), (266.4, -29.0), (10.0, 0.0), (359.9, 75.0)])
def test_tan_product_vs_oracle_and_analytic(skydir):
    g = WcsGeom.create(skydir=skydir, binsz=0.05, width=(6.0, 4.0), proj="TAN", frame="icrs")
    o = ImageGeom.create(skydir=skydir, binsz=0.05, npix=(120, 80), proj="TAN")
    assert g.npix == (120, 80) and g.crval == skydir and g.crpix == (60.5, 40.5)
    c = g.get_coord()
    lon, lat = o.get_coord()
    dlon = (c["lon"] - lon + 180.0) % 360.0 - 180.0
    assert np.abs(dlon).max() < 1e-11 and np.abs(c["lat"] - lat).max() < 1e-12
    yy, xx = np.mgrid[0:80, 0:120]
    px, py = g.coord_to_pix((c["lon"], c["lat"]))
    assert np.abs(px - xx).max() < 1e-10 and np.abs(py - yy).max() < 1e-10
    opx, opy = o.world_to_pix(lon, lat)
    assert np.abs(opx - xx).max() < 1e-8 and np.abs(opy - yy).max() < 1e-8
    # gnomonic: a pixel R degrees from the reference pixel (in the plane) is atan(R) from the reference point on the sky
    sep = angular_separation(c["lon"] * DEG, c["lat"] * DEG, skydir[0] * DEG, skydir[1] * DEG) / DEG
    r_plane = 0.05 * np.hypot(xx - 59.5, yy - 39.5)
    np.testing.assert_allclose(sep, np.degrees(np.arctan(np.radians(r_plane))), atol=1e-12)
    # north is up, east is left
    assert c["lat"][-1, 60] > c["lat"][0, 60]
    assert ((c["lon"][40, 0] - c["lon"][40, -1] + 180.0) % 360.0 - 180.0) > 0
    # solid angle: pixels shrink away from the tangent point as cos^3(sep)
    omega = g.solid_angle()
    np.testing.assert_allclose(omega, o.solid_angle(), rtol=1e-9)
    np.testing.assert_allclose(omega / omega[39:41, 59:61].mean(), np.cos(sep * DEG) ** 3, rtol=2e-3)
    # cutouts follow the same integer rules as on a CAR geometry
    ys, xs = g.cutout_slices_for((c["lon"][30, 50], c["lat"][30, 50]), 1.0, odd_npix=True)
    _, (oys, oxs) = o.cutout_info((lon[30, 50], lat[30, 50]), 1.0, odd_npix=True)
    assert (ys, xs) == (oys, oxs) and xs.stop - xs.start == 21 and ys == slice(20, 41)


def test_tan_matches_car_at_the_equator_to_third_order():
    car = WcsGeom.create(skydir=(30.0, 0.0), binsz=0.02, width=(2.0, 2.0), proj="CAR", frame="galactic")
    tan = WcsGeom.create(skydir=(30.0, 0.0), binsz=0.02, width=(2.0, 2.0), proj="TAN", frame="galactic")
    a, b = car.get_coord(), tan.get_coord()
    assert np.abs(a["lon"] - b["lon"]).max() < 4e-4 and np.abs(a["lat"] - b["lat"]).max() < 4e-4  # third order in the 1.4 deg to the corner
    assert car != tan
