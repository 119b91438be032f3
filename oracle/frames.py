"""Oracle: ICRS <-> Galactic coordinates, for spatial models whose frame differs from the map's.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference evaluates such a model on ``geom.get_coord(frame=model.frame)`` (modeling/models/spatial.py:196-220) and
sends ``model.position`` through ``skycoord_to_pixel`` for cutouts (maps/wcs/geom.py:886-930); both go through astropy's
frame graph.  astropy (>=6.1.5,<8.0 per the reference's pyproject.toml:23-35) is not importable here, so its published
definition is restated -- in the classical spherical-trigonometry form, not the matrix product the product code uses:

* ICRS -> FK5 (J2000): the frame-bias rotation of USNO circular 179, eta0 = -19.9 mas about x, xi0 = 9.1 mas about y,
  da0 = -22.9 mas about z (coordinates/builtin_frames/icrs_fk5_transforms.py);
* FK5 (J2000) -> Galactic: north galactic pole (alpha_G, delta_G) = (192.8594812065348, 27.12825118085622) deg and
  galactic longitude of the north celestial pole l_NCP = 122.9319185680026 deg (coordinates/builtin_frames/galactic.py):
      sin b               = sin d sin d_G + cos d cos d_G cos(a - a_G)
      cos b sin(l_NCP - l) = cos d sin(a - a_G)
      cos b cos(l_NCP - l) = sin d cos d_G - cos d sin d_G cos(a - a_G)

**Parity unpinned** beyond the ICRS positions of the galactic centre and pole astropy is known to print
(266.40498829, -28.93617776) and (192.85947789, 27.12825241), which the tests check.
"""
import numpy as np

DEG = np.pi / 180.0
MAS = DEG / 3600000.0
ETA0, XI0, DA0 = -19.9 * MAS, 9.1 * MAS, -22.9 * MAS
ALPHA_G, DELTA_G, L_NCP = 192.8594812065348 * DEG, 27.12825118085622 * DEG, 122.9319185680026 * DEG


def _vec(lon, lat):
    return np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)


def _lonlat(x, y, z):
    return np.arctan2(y, x), np.arctan2(z, np.hypot(x, y))


def _bias(x, y, z, inverse=False):
    """ICRS -> FK5 J2000: Rx(-eta0) Ry(xi0) Rz(da0) applied to the vector (passive rotations), or its inverse."""
    def rz(a, v):
        return np.cos(a) * v[0] + np.sin(a) * v[1], -np.sin(a) * v[0] + np.cos(a) * v[1], v[2]

    def ry(a, v):
        return np.cos(a) * v[0] - np.sin(a) * v[2], v[1], np.sin(a) * v[0] + np.cos(a) * v[2]

    def rx(a, v):
        return v[0], np.cos(a) * v[1] + np.sin(a) * v[2], -np.sin(a) * v[1] + np.cos(a) * v[2]

    v = (x, y, z)
    if not inverse:
        return rx(-ETA0, ry(XI0, rz(DA0, v)))
    return rz(-DA0, ry(-XI0, rx(ETA0, v)))


def icrs_to_galactic(ra, dec):
    """deg -> deg."""
    a, d = _lonlat(*_bias(*_vec(np.asarray(ra, float) * DEG, np.asarray(dec, float) * DEG)))
    sin_b = np.sin(d) * np.sin(DELTA_G) + np.cos(d) * np.cos(DELTA_G) * np.cos(a - ALPHA_G)
    y = np.cos(d) * np.sin(a - ALPHA_G)
    x = np.sin(d) * np.cos(DELTA_G) - np.cos(d) * np.sin(DELTA_G) * np.cos(a - ALPHA_G)
    l = L_NCP - np.arctan2(y, x)
    b = np.arctan2(sin_b, np.hypot(x, y))
    return (l / DEG) % 360.0, b / DEG


def galactic_to_icrs(l, b):
    """deg -> deg."""
    l, b = np.asarray(l, float) * DEG, np.asarray(b, float) * DEG
    sin_d = np.sin(b) * np.sin(DELTA_G) + np.cos(b) * np.cos(DELTA_G) * np.cos(L_NCP - l)
    y = np.cos(b) * np.sin(L_NCP - l)
    x = np.sin(b) * np.cos(DELTA_G) - np.cos(b) * np.sin(DELTA_G) * np.cos(L_NCP - l)
    a = ALPHA_G + np.arctan2(y, x)
    d = np.arctan2(sin_d, np.hypot(x, y))
    ra, dec = _lonlat(*_bias(*_vec(a, d), inverse=True))
    return (ra / DEG) % 360.0, dec / DEG


def transform(lon, lat, src, dst):
    """(lon, lat) [deg] of frame ``src`` in frame ``dst`` ("icrs" / "galactic"; None = unspecified = no transform)."""
    if src is None or dst is None or src == dst:
        return np.asarray(lon, float), np.asarray(lat, float)
    if (src, dst) == ("icrs", "galactic"):
        return icrs_to_galactic(lon, lat)
    if (src, dst) == ("galactic", "icrs"):
        return galactic_to_icrs(lon, lat)
    raise NotImplementedError(f"oracle: frame transform {src!r} -> {dst!r}")
