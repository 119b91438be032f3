"""Small host-side helpers: unit-string parsing (astropy is not a dependency), sky coordinates, RNG."""
import numbers
import re

import numpy as np

DEG = np.pi / 180.0

_ENERGY_TO_TEV = {"ev": 1e-12, "kev": 1e-9, "mev": 1e-6, "gev": 1e-3, "tev": 1.0, "pev": 1e3}
_ANGLE_TO_DEG = {"deg": 1.0, "rad": 180.0 / np.pi, "arcmin": 1.0 / 60.0, "arcsec": 1.0 / 3600.0}
# The package's fixed unit system (include/gpy_b200.h): units are labels, not conversions, so every other spelling of
# a flux / inverse-energy / exposure unit is converted here or refused -- never kept silently.
_FLUX_TO_CANONICAL = {  # -> cm-2 s-1 TeV-1
    "cm-2 s-1 tev-1": 1.0, "1 / (cm2 s tev)": 1.0, "cm-2 s-1 gev-1": 1e3, "cm-2 s-1 mev-1": 1e6,
    "cm-2 s-1 kev-1": 1e9, "m-2 s-1 tev-1": 1e-4, "m-2 s-1 gev-1": 1e-1, "m-2 s-1 mev-1": 1e2,
}
_INVERSE_ENERGY_TO_CANONICAL = {"tev-1": 1.0, "1 / tev": 1.0, "gev-1": 1e3, "mev-1": 1e6, "kev-1": 1e9, "pev-1": 1e-3}
_PASS_THROUGH_UNITS = {"", "tev", "deg", "m2 s", "cm2 s", "sr-1", "1 / sr", "sr", "s", "m2"}
_NUM = re.compile(r"^\s*([-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|nan|inf|-inf)\s*(.*)$")


def parse_quantity(value, kind=None):
    """Return (float value, unit string) from a number or a string such as ``"0.2 deg"``.

    ``kind`` = "energy" converts to TeV, "angle" to deg; other units are kept verbatim.  The fixed unit
    system of this package is TeV / deg / cm-2 s-1 TeV-1 / m2 s (see gpy_b200.h).
    """
    unit = ""
    if hasattr(value, "unit") and hasattr(value, "value"):  # astropy Quantity, if the user has astropy
        unit = str(value.unit)
        value = value.value
    if isinstance(value, str):
        m = _NUM.match(value)
        if not m:
            raise ValueError(f"Cannot parse quantity {value!r}")
        value, unit = float(m.group(1)), m.group(2).strip()
    elif isinstance(value, numbers.Number):
        value = float(value)
    else:
        value = np.asarray(value, dtype=float)
    key = unit.lower()
    if kind == "energy" and key:
        if key not in _ENERGY_TO_TEV:
            raise ValueError(f"Not an energy unit: {unit!r}")
        return value * _ENERGY_TO_TEV[key], "TeV"
    if kind == "angle" and key:
        if key not in _ANGLE_TO_DEG:
            raise ValueError(f"Not an angle unit: {unit!r}")
        return value * _ANGLE_TO_DEG[key], "deg"
    if key in _ENERGY_TO_TEV and key != "tev":
        return value * _ENERGY_TO_TEV[key], "TeV"
    if key in _ANGLE_TO_DEG and key != "deg":
        return value * _ANGLE_TO_DEG[key], "deg"
    norm = " ".join(key.split())
    if norm in _FLUX_TO_CANONICAL:
        return value * _FLUX_TO_CANONICAL[norm], "cm-2 s-1 TeV-1"
    if norm in _INVERSE_ENERGY_TO_CANONICAL:
        return value * _INVERSE_ENERGY_TO_CANONICAL[norm], "TeV-1"
    if norm == "cm2 s":
        return value * 1e-4, "m2 s"
    if norm not in _PASS_THROUGH_UNITS:
        raise ValueError(f"Unit {unit!r} is not one this package converts (TeV, deg, cm-2 s-1 TeV-1, TeV-1, m2 s and "
                         "their common multiples): refusing to keep it as a label")
    return value, unit


def to_value(value, kind=None):
    return parse_quantity(value, kind)[0]


def angular_separation(lon1, lat1, lon2, lat2):
    """Great-circle separation (Vincenty), radians in / radians out."""
    sdlon = np.sin(lon2 - lon1)
    cdlon = np.cos(lon2 - lon1)
    slat1, slat2 = np.sin(lat1), np.sin(lat2)
    clat1, clat2 = np.cos(lat1), np.cos(lat2)
    num1 = clat2 * sdlon
    num2 = clat1 * slat2 - slat1 * clat2 * cdlon
    den = slat1 * slat2 + clat1 * clat2 * cdlon
    return np.arctan2(np.hypot(num1, num2), den)


def position_angle(lon1, lat1, lon2, lat2):
    deltalon = lon2 - lon1
    colat = np.cos(lat2)
    x = np.sin(lat2) * np.cos(lat1) - colat * np.sin(lat1) * np.cos(deltalon)
    y = np.sin(deltalon) * colat
    return np.arctan2(y, x)


# -- celestial frames ----------------------------------------------------------------------------------------------
# ICRS <-> Galactic as astropy does it (coordinates/builtin_frames: icrs_fk5_transforms.py, galactic_transforms.py,
# galactic.py): ICRS -> FK5(J2000) by the frame-bias matrix of USNO circular 179 (eta0 = -19.9 mas, xi0 = 9.1 mas,
# da0 = -22.9 mas), FK5(J2000) -> Galactic by the rotation that puts the north galactic pole
# (ra, dec) = (192.8594812065348, 27.12825118085622) deg on the z axis with the longitude zero point
# lon0 = 122.9319185680026 deg.  rotation_matrix(angle, axis) there is the passive rotation (coordinates of a fixed
# vector in the rotated basis).  The numbers are restated from the published astropy source (astropy itself cannot be
# imported here): parity unpinned beyond the positions of the galactic centre and pole checked in the tests.
FRAMES = ("galactic", "icrs")


def _rot(angle_deg, axis):
    a = np.radians(angle_deg)
    c, s = np.cos(a), np.sin(a)
    if axis == "x":
        return np.array([[1.0, 0.0, 0.0], [0.0, c, s], [0.0, -s, c]])
    if axis == "y":
        return np.array([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]])
    return np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])


_ICRS_TO_FK5 = _rot(19.9 / 3600000.0, "x") @ _rot(9.1 / 3600000.0, "y") @ _rot(-22.9 / 3600000.0, "z")
_FK5_TO_GAL = _rot(180.0 - 122.9319185680026, "z") @ _rot(90.0 - 27.12825118085622, "y") @ _rot(192.8594812065348, "z")
ICRS_TO_GALACTIC = _FK5_TO_GAL @ _ICRS_TO_FK5


def frame_matrix(src, dst):
    """3 x 3 matrix taking unit vectors of frame ``src`` to frame ``dst`` (None when they are the same frame)."""
    if src == dst:
        return None
    if (src, dst) == ("icrs", "galactic"):
        return ICRS_TO_GALACTIC
    if (src, dst) == ("galactic", "icrs"):
        return ICRS_TO_GALACTIC.T
    raise NotImplementedError(f"frame transform {src!r} -> {dst!r} (frames: {FRAMES})")


def transform_lonlat(lon, lat, src, dst):
    """(lon, lat) in deg from frame ``src`` to frame ``dst``; lon in [0, 360)."""
    m = frame_matrix(src, dst)
    lon, lat = np.asarray(lon, dtype=float), np.asarray(lat, dtype=float)
    if m is None:
        return lon, lat
    lo, la = lon * DEG, lat * DEG
    x, y, z = np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.sin(la)
    x2 = m[0, 0] * x + m[0, 1] * y + m[0, 2] * z
    y2 = m[1, 0] * x + m[1, 1] * y + m[1, 2] * z
    z2 = m[2, 0] * x + m[2, 1] * y + m[2, 2] * z
    return np.arctan2(y2, x2) / DEG % 360.0, np.arctan2(z2, np.hypot(x2, y2)) / DEG


class SkyCoord:
    """Minimal (lon, lat) in degrees with a frame label; enough for map centres and model positions."""

    def __init__(self, lon, lat, unit="deg", frame="icrs"):
        def ang(v):
            if isinstance(v, str) or hasattr(v, "unit"):
                return to_value(v, "angle")
            return np.asarray(v, dtype=float) * _ANGLE_TO_DEG[unit]

        self.lon = ang(lon)
        self.lat = ang(lat)
        self.frame = frame

    l = ra = property(lambda self: self.lon)
    b = dec = property(lambda self: self.lat)

    def transform_to(self, frame):
        if frame == self.frame:
            return self
        lon, lat = transform_lonlat(self.lon, self.lat, self.frame, frame)
        return SkyCoord(lon, lat, frame=frame)

    galactic = property(lambda self: self.transform_to("galactic"))
    icrs = property(lambda self: self.transform_to("icrs"))

    def separation(self, other):
        """Separation in deg (``other`` is brought into this frame first, as astropy does)."""
        if isinstance(other, SkyCoord) and other.frame != self.frame:
            other = other.transform_to(self.frame)
        return angular_separation(self.lon * DEG, self.lat * DEG, np.asarray(other.lon) * DEG,
                                  np.asarray(other.lat) * DEG) / DEG

    def __iter__(self):
        return iter((self.lon, self.lat))

    def __repr__(self):
        return f"SkyCoord(lon={self.lon}, lat={self.lat}, frame={self.frame!r})"


def as_lonlat(skydir):
    """(lon, lat) in deg from a tuple or SkyCoord-like."""
    if skydir is None:
        return 0.0, 0.0
    if isinstance(skydir, SkyCoord):
        return float(skydir.lon), float(skydir.lat)
    if hasattr(skydir, "spherical"):  # astropy SkyCoord
        return float(skydir.spherical.lon.deg), float(skydir.spherical.lat.deg)
    lon, lat = skydir
    return to_value(lon, "angle"), to_value(lat, "angle")


def get_random_state(init):
    """utils/random/utils.py:81-92 of the reference: legacy MT19937 RandomState."""
    if isinstance(init, (numbers.Integral, np.integer)):
        return np.random.RandomState(int(init))
    if init == "random-seed":
        return np.random.RandomState(None)
    if init == "global-rng":
        return np.random.mtrand._rand
    if isinstance(init, np.random.RandomState):
        return init
    raise ValueError(f"{init} cannot be used to seed a numpy.random.RandomState instance.")
