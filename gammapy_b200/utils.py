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

    def separation(self, other):
        """Separation in deg."""
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
