"""Oracle: WCS image geometry (CAR, TAN), spherical trig, cutouts.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, float-for-float where the arithmetic is under /root/reference and
from the published formulas where it lives in astropy/wcslib
(astropy >=6.1.5,<8.0 per the reference's pyproject.toml:23-35):

* ``WcsGeom.create``            maps/wcs/geom.py:294-417   (crpix=(n+1)/2, cdelt=(-binsz,+binsz))
* ``WcsGeom.upsample``          maps/wcs/geom.py:766-779   (+ get_resampled_wcs: crpix'=(crpix-0.5)*f+0.5)
* ``WcsGeom.solid_angle``       maps/wcs/geom.py:803-854   (planar-triangle formula on pixel corners)
* ``WcsGeom.cutout``            maps/wcs/geom.py:886-930   (astropy Cutout2D / overlap_slices rule)
* ``WcsGeom.cutout_slices``     maps/wcs/geom.py:168-193
* ``WcsGeom.to_odd_npix``       maps/wcs/geom.py:1106-1138
* ``round_up_to_odd``           utils/array.py:110
* astropy ``angular_separation`` (Vincenty) and ``position_angle``.

CAR with ``crval = (lon0, 0)`` is the plain linear map
``lon = lon0 + cdelt_x (i + 1 - crpix_x)``, ``lat = cdelt_y (j + 1 - crpix_y)``
(wcslib: native pole at delta_p = 90 deg reduces the spherical rotation to a
longitude offset).  CAR about a reference point off the equator -- what
``WcsGeom.create(skydir=(lon, lat != 0))`` builds, geom.py:370-410 -- is an oblique
plate carree: restated from FITS WCS paper II in its general form (native pole from
eqs. 8-10 with the default LONPOLE / LATPOLE, rotations eqs. 2 and 5), parity
unpinned like TAN (wcslib cannot be run here).
"""
import numpy as np

DEG = np.pi / 180.0  # astropy deg->rad factor 0.017453292519943295


class NoOverlapError(ValueError):
    """astropy.nddata.utils.NoOverlapError stand-in."""


def angular_separation(lon1, lat1, lon2, lat2):
    """Vincenty great-circle separation, all angles in radians (astropy.coordinates.angular_separation)."""
    sdlon = np.sin(lon2 - lon1)
    cdlon = np.cos(lon2 - lon1)
    slat1 = np.sin(lat1)
    slat2 = np.sin(lat2)
    clat1 = np.cos(lat1)
    clat2 = np.cos(lat2)
    num1 = clat2 * sdlon
    num2 = clat1 * slat2 - slat1 * clat2 * cdlon
    denominator = slat1 * slat2 + clat1 * clat2 * cdlon
    return np.arctan2(np.hypot(num1, num2), denominator)


def position_angle(lon1, lat1, lon2, lat2):
    """Position angle (radians, East of North) from point 1 to point 2 (astropy.coordinates.position_angle)."""
    deltalon = lon2 - lon1
    colat = np.cos(lat2)
    x = np.sin(lat2) * np.cos(lat1) - colat * np.sin(lat1) * np.cos(deltalon)
    y = np.sin(deltalon) * colat
    return np.arctan2(y, x)


def round_up_to_odd(f):
    """utils/array.py:110."""
    return (np.ceil(f) // 2 * 2 + 1).astype(int)


def overlap_slices(large_shape, small_shape, position, mode="trim"):
    """astropy.nddata.utils.overlap_slices for modes 'trim' and 'partial' (2-D, (y, x) order)."""
    if any(not np.isfinite(p) for p in position):
        raise ValueError("Input position contains invalid values (NaNs or infs).")
    edges_min = [int(np.ceil(pos - (s / 2.0))) for pos, s in zip(position, small_shape)]
    edges_max = [int(np.ceil(pos + (s / 2.0))) for pos, s in zip(position, small_shape)]
    for e_max in edges_max:
        if e_max <= 0:
            raise NoOverlapError("Arrays do not overlap.")
    for e_min, large in zip(edges_min, large_shape):
        if e_min >= large:
            raise NoOverlapError("Arrays do not overlap.")
    slices_large = tuple(
        slice(max(0, e_min), min(large, e_max))
        for e_min, e_max, large in zip(edges_min, edges_max, large_shape)
    )
    if mode == "trim":
        slices_small = tuple(slice(0, s.stop - s.start) for s in slices_large)
    else:
        slices_small = tuple(
            slice(max(0, -e_min), min(large - e_min, e_max - e_min))
            for e_min, e_max, large in zip(edges_min, edges_max, large_shape)
        )
    return slices_large, slices_small


class ImageGeom:
    """2-D WCS image geometry; (nx, ny) pixels of ``binsz`` deg, 0-based pixel coords.  ``proj``: "CAR" (about the
    equator the separable linear map; about any other reference point the oblique plate carree, see ``_car_pole``), or
    "TAN" about (crval_lon, crval_lat), restated from the FITS WCS paper II (Calabretta & Greisen 2002) in its
    native-spherical-coordinate form: eq. (54) R = (180/pi) cot(theta), eq. (14)-(15) x = R sin(phi), y = -R cos(phi), and
    the spherical rotation eq. (2) / (5) with the native pole at the reference point and LONPOLE = 180 deg.  wcslib
    itself cannot be run here: **parity unpinned** for TAN beyond the analytic checks of tests/test_host_logic.py."""

    def __init__(self, nx, ny, binsz, crval_lon, crpix_x, crpix_y, crval_lat=0.0, proj="CAR"):
        if proj not in ("CAR", "TAN"):
            raise NotImplementedError(f"oracle: projection {proj!r}")
        if proj == "CAR" and not -90.0 < crval_lat < 90.0:
            raise ValueError("oracle: CAR reference latitude must lie inside (-90, 90) deg")
        self.proj = proj
        self.nx = int(nx)
        self.ny = int(ny)
        self.binsz = float(binsz)
        self.crval_lon = float(crval_lon)
        self.crval_lat = float(crval_lat)
        self.crpix_x = float(crpix_x)
        self.crpix_y = float(crpix_y)

    # -- construction ---------------------------------------------------------------------
    @classmethod
    def create(cls, skydir=(0.0, 0.0), binsz=0.5, width=None, npix=None, proj="CAR"):
        """maps/wcs/geom.py:294-417 (square pixels only)."""
        if npix is None:
            if not isinstance(width, (tuple, list)):
                width = (width, width)
            npix = (
                int(np.rint(float(width[0]) / binsz)),
                int(np.rint(float(width[1]) / binsz)),
            )
        elif not isinstance(npix, (tuple, list)):
            npix = (int(npix), int(npix))
        nx, ny = int(npix[0]), int(npix[1])
        return cls(nx, ny, binsz, skydir[0], (nx + 1) / 2.0, (ny + 1) / 2.0, skydir[1], proj=proj)

    def copy(self, **kw):
        d = dict(nx=self.nx, ny=self.ny, binsz=self.binsz, crval_lon=self.crval_lon,
                 crpix_x=self.crpix_x, crpix_y=self.crpix_y, crval_lat=self.crval_lat, proj=self.proj)
        d.update(kw)
        return ImageGeom(**d)

    @property
    def data_shape(self):
        return (self.ny, self.nx)

    @property
    def width(self):
        """maps/wcs/geom.py:223-227: cdelt * npix (lon, lat) in deg."""
        return (self.binsz * self.nx, self.binsz * self.ny)

    @property
    def pixel_scale(self):
        """np.max(pixel_scales) – proj_plane_pixel_scales = |cdelt| for a diagonal PC matrix."""
        return self.binsz

    @property
    def center_pix(self):
        return ((self.nx - 1) / 2.0, (self.ny - 1) / 2.0)

    @property
    def center_skydir(self):
        lon, lat = self.pix_to_world(*self.center_pix)
        return float(lon), float(lat)

    # -- transforms ------------------------------------------------------------------------
    def _car_pole(self):
        """(alpha_p, delta_p, phi_p) in radians of an oblique plate carree (theta_0 = phi_0 = 0), FITS WCS paper II
        section 2.4: LONPOLE defaults to 0 deg when delta_0 >= theta_0 and to 180 deg otherwise; eq. (8)
        delta_p = arg(cos(theta_0) cos(phi_p - phi_0), sin(theta_0)) +- acos(sin(delta_0) / sqrt(1 - cos^2(theta_0)
        sin^2(phi_p - phi_0))), the solution nearer LATPOLE (default +90 deg) when both are valid; eqs. (9), (10)
        sin(alpha_0 - alpha_p) = sin(phi_p - phi_0) cos(theta_0) / cos(delta_0),
        cos(alpha_0 - alpha_p) = (sin(theta_0) - sin(delta_p) sin(delta_0)) / (cos(delta_p) cos(delta_0))."""
        delta_0, theta_0, phi_0 = self.crval_lat * DEG, 0.0, 0.0
        phi_p = 0.0 if delta_0 >= theta_0 else np.pi
        u = np.arctan2(np.sin(theta_0), np.cos(theta_0) * np.cos(phi_p - phi_0))
        v = np.arccos(np.sin(delta_0) / np.sqrt(1.0 - (np.cos(theta_0) * np.sin(phi_p - phi_0)) ** 2))
        candidates = []
        for cand in (u + v, u - v):
            cand = (cand + np.pi) % (2 * np.pi) - np.pi        # into [-180, 180) deg
            if -np.pi / 2 - 1e-12 <= cand <= np.pi / 2 + 1e-12:
                candidates.append(cand)
        delta_p = min(candidates, key=lambda c: abs(np.pi / 2 - c))
        s = np.sin(phi_p - phi_0) * np.cos(theta_0) / np.cos(delta_0)
        c = (np.sin(theta_0) - np.sin(delta_p) * np.sin(delta_0)) / (np.cos(delta_p) * np.cos(delta_0))
        alpha_p = self.crval_lon * DEG - np.arctan2(s, c)
        return alpha_p, delta_p, phi_p

    def pix_to_world(self, px, py):
        """wcs_pix2world(px, py, 0) in deg."""
        px = np.asarray(px, dtype=float)
        py = np.asarray(py, dtype=float)
        if self.proj == "TAN":
            x = (-self.binsz) * (px + 1.0 - self.crpix_x)
            y = self.binsz * (py + 1.0 - self.crpix_y)
            x, y = np.broadcast_arrays(x, y)
            r_theta = np.sqrt(x * x + y * y)
            phi = np.arctan2(x, -y)                            # eq. (14), (15)
            theta = np.arctan2(180.0 / np.pi, r_theta)         # eq. (55)
            phi_p, delta_p = np.pi, self.crval_lat * DEG
            delta = np.arcsin(np.clip(np.sin(theta) * np.sin(delta_p)
                                      + np.cos(theta) * np.cos(delta_p) * np.cos(phi - phi_p), -1, 1))   # eq. (2)
            alpha = self.crval_lon * DEG + np.arctan2(
                -np.cos(theta) * np.sin(phi - phi_p),
                np.sin(theta) * np.cos(delta_p) - np.cos(theta) * np.sin(delta_p) * np.cos(phi - phi_p))
            at_ref = r_theta == 0
            return (np.where(at_ref, self.crval_lon, alpha / DEG), np.where(at_ref, self.crval_lat, delta / DEG))
        if self.crval_lat != 0.0:   # oblique plate carree: eqs. (83), (84) phi = x, theta = y, then the rotation eq. (2)
            phi = (-self.binsz) * (px + 1.0 - self.crpix_x) * DEG
            theta = self.binsz * (py + 1.0 - self.crpix_y) * DEG
            phi, theta = np.broadcast_arrays(phi, theta)
            alpha_p, delta_p, phi_p = self._car_pole()
            delta = np.arcsin(np.clip(np.sin(theta) * np.sin(delta_p)
                                      + np.cos(theta) * np.cos(delta_p) * np.cos(phi - phi_p), -1, 1))
            alpha = alpha_p + np.arctan2(
                -np.cos(theta) * np.sin(phi - phi_p),
                np.sin(theta) * np.cos(delta_p) - np.cos(theta) * np.sin(delta_p) * np.cos(phi - phi_p))
            # the celestial longitude on the branch next to the reference longitude (a map does not straddle +-180 deg of it)
            lon = self.crval_lon + ((alpha / DEG - self.crval_lon + 180.0) % 360.0 - 180.0)
            return lon, delta / DEG
        lon = self.crval_lon + (-self.binsz) * (px + 1.0 - self.crpix_x)
        lat = self.binsz * (py + 1.0 - self.crpix_y)
        lon, lat = np.broadcast_arrays(lon, lat)
        return lon, lat

    def world_to_pix(self, lon, lat):
        """wcs_world2pix(lon, lat, 0); longitude difference wrapped into [-180, 180)."""
        if self.proj == "TAN":
            alpha, delta = np.asarray(lon, dtype=float) * DEG, np.asarray(lat, dtype=float) * DEG
            alpha_p, delta_p, phi_p = self.crval_lon * DEG, self.crval_lat * DEG, np.pi
            phi = phi_p + np.arctan2(-np.cos(delta) * np.sin(alpha - alpha_p),
                                     np.sin(delta) * np.cos(delta_p) - np.cos(delta) * np.sin(delta_p) * np.cos(alpha - alpha_p))  # eq. (5)
            theta = np.arcsin(np.clip(np.sin(delta) * np.sin(delta_p) + np.cos(delta) * np.cos(delta_p) * np.cos(alpha - alpha_p), -1, 1))
            with np.errstate(divide="ignore", invalid="ignore"):
                r_theta = np.where(theta > 0, (180.0 / np.pi) / np.tan(theta), np.nan)   # eq. (54)
            x, y = r_theta * np.sin(phi), -r_theta * np.cos(phi)
            return x / (-self.binsz) + self.crpix_x - 1.0, y / self.binsz + self.crpix_y - 1.0
        if self.crval_lat != 0.0:   # oblique plate carree: inverse rotation eq. (5), then x = phi, y = theta
            alpha, delta = np.asarray(lon, dtype=float) * DEG, np.asarray(lat, dtype=float) * DEG
            alpha_p, delta_p, phi_p = self._car_pole()
            phi = phi_p + np.arctan2(-np.cos(delta) * np.sin(alpha - alpha_p),
                                     np.sin(delta) * np.cos(delta_p) - np.cos(delta) * np.sin(delta_p) * np.cos(alpha - alpha_p))
            theta = np.arcsin(np.clip(np.sin(delta) * np.sin(delta_p) + np.cos(delta) * np.cos(delta_p) * np.cos(alpha - alpha_p), -1, 1))
            phi = (phi + np.pi) % (2 * np.pi) - np.pi
            return (phi / DEG) / (-self.binsz) + self.crpix_x - 1.0, (theta / DEG) / self.binsz + self.crpix_y - 1.0
        dlon = (np.asarray(lon, dtype=float) - self.crval_lon + 180.0) % 360.0 - 180.0
        px = dlon / (-self.binsz) + self.crpix_x - 1.0
        py = np.asarray(lat, dtype=float) / self.binsz + self.crpix_y - 1.0
        return px, py

    def get_coord(self, mode="center"):
        """Pixel-centre (or edge) lon/lat grids in deg, shape (ny[+1], nx[+1])."""
        if mode == "edges":
            x = np.arange(-0.5, self.nx, dtype=float)
            y = np.arange(-0.5, self.ny, dtype=float)
        else:
            x = np.arange(self.nx, dtype=float)
            y = np.arange(self.ny, dtype=float)
        return self.pix_to_world(x[np.newaxis, :], y[:, np.newaxis])

    # -- derived geometries ------------------------------------------------------------------
    def upsample(self, factor):
        """maps/wcs/geom.py:766-771 + maps/wcs/geom.py get_resampled_wcs (crpix'=(crpix-0.5)*f+0.5, cdelt/=f)."""
        return ImageGeom(
            self.nx * factor, self.ny * factor, self.binsz / factor, self.crval_lon,
            (self.crpix_x - 0.5) * factor + 0.5, (self.crpix_y - 0.5) * factor + 0.5, self.crval_lat, proj=self.proj,
        )

    def to_odd_npix(self, max_radius=None):
        """maps/wcs/geom.py:1106-1138."""
        if max_radius is None:
            width = max(self.width)
        else:
            width = 2 * max_radius
        width_npix = width / self.binsz
        npix = int(round_up_to_odd(np.float64(width_npix)))
        return ImageGeom.create(skydir=self.center_skydir, binsz=self.binsz, npix=npix, proj=self.proj)

    def cutout_info(self, position, width, mode="trim", odd_npix=False, min_npix=1):
        """maps/wcs/geom.py:886-930 -> (cutout geom, parent (y, x) slices).

        ``position`` = (lon, lat) deg; ``width`` = float or (lon_width, lat_width) deg.
        """
        if not isinstance(width, (tuple, list, np.ndarray)):
            width = (width, width)
        width = np.array([float(width[0]), float(width[1])])
        width_npix = np.clip(width / self.binsz, min_npix, None)
        if odd_npix:
            width_npix = round_up_to_odd(width_npix)
        # Cutout2D: size in (ny, nx) order, shape = int(np.round(size))
        shape = (int(np.round(width_npix[1])), int(np.round(width_npix[0])))
        px, py = self.world_to_pix(position[0], position[1])
        slices_large, _ = overlap_slices(self.data_shape, shape, (float(py), float(px)), mode=mode)
        ys, xs = slices_large
        geom = ImageGeom(
            xs.stop - xs.start, ys.stop - ys.start, self.binsz, self.crval_lon,
            self.crpix_x - xs.start, self.crpix_y - ys.start, self.crval_lat, proj=self.proj,
        )
        return geom, (ys, xs)

    def cutout(self, position, width, **kw):
        return self.cutout_info(position, width, **kw)[0]

    def cutout_slices(self, parent, mode="partial"):
        """maps/wcs/geom.py:168-193: slices of this (cutout) geom inside ``parent``."""
        px, py = parent.world_to_pix(*self.center_skydir)
        large, small = overlap_slices(parent.data_shape, self.data_shape, (float(py), float(px)), mode=mode)
        return {"parent-slices": large, "cutout-slices": small}

    # -- solid angle ---------------------------------------------------------------------------
    def solid_angle(self):
        """maps/wcs/geom.py:816-854, sr, shape (ny, nx)."""
        lon, lat = self.get_coord(mode="edges")
        lon = lon * DEG
        lat = lat * DEG

        def pt(sy, sx):
            return lon[sy, sx], lat[sy, sx]

        lo, hi = slice(None, -1), slice(1, None)
        low_left = pt(lo, lo)
        low_right = pt(hi, lo)
        up_left = pt(lo, hi)
        up_right = pt(hi, hi)

        low = angular_separation(*low_left, *low_right)
        left = angular_separation(*low_left, *up_left)
        up = angular_separation(*up_left, *up_right)
        right = angular_separation(*low_right, *up_right)

        angle_low_right = position_angle(*low_right, *up_right) - position_angle(*low_right, *low_left)
        angle_up_left = position_angle(*up_left, *up_right) - position_angle(*low_left, *up_left)

        area_low_right = 0.5 * low * right * np.sin(angle_low_right)
        area_up_left = 0.5 * up * left * np.sin(angle_up_left)
        return np.abs(area_low_right + area_up_left)
