"""Host-side map containers mirroring the subset of ``gammapy.maps`` the likelihood path touches.

Reference: maps/axes.py (MapAxis), maps/wcs/geom.py (WcsGeom), maps/core.py + maps/wcs/ndmap.py (Map).
Geometries are CAR (plate carree about any reference point) or TAN images with square pixels (the projections the
CUDA kernels implement); data live in numpy arrays on the host and are mirrored to HBM by ``MapDataset``.
"""
import copy as _copy

import numpy as np

from .utils import DEG, SkyCoord, angular_separation, as_lonlat, position_angle, to_value

__all__ = ["MapAxis", "WcsGeom", "Map", "WcsNDMap"]


class MapAxis:
    """Non-spatial axis (maps/axes.py:85).  Energy axes are log-interpolated, others linear."""

    def __init__(self, nodes, interp="lin", name="", node_type="edges", unit=""):
        self.name = name
        self.unit = unit
        self.interp = interp
        self.node_type = node_type
        self._nodes = np.asarray(nodes, dtype=float)
        if self._nodes.ndim != 1 or len(self._nodes) < (2 if node_type == "edges" else 1):
            raise ValueError("MapAxis: need a 1D array of nodes")
        if node_type == "edges":
            self._nbin = len(self._nodes) - 1
            self._pix_offset = -0.5
        else:
            self._nbin = len(self._nodes)
            self._pix_offset = 0.0

    # -- constructors -----------------------------------------------------------------------
    @classmethod
    def from_energy_bounds(cls, energy_min, energy_max, nbin, unit=None, per_decade=False, name="energy",
                           node_type="edges"):
        """maps/axes.py from_energy_bounds: geomspace edges."""
        emin = to_value(energy_min if unit is None or isinstance(energy_min, str) else f"{energy_min} {unit}", "energy")
        emax = to_value(energy_max if unit is None or isinstance(energy_max, str) else f"{energy_max} {unit}", "energy")
        if per_decade:
            nbin = np.ceil(np.log10(emax / emin) * nbin)
        nodes = np.geomspace(emin, emax, int(nbin) + 1)
        return cls(nodes, interp="log", name=name, node_type=node_type, unit="TeV")

    @classmethod
    def from_energy_edges(cls, energy_edges, unit=None, name="energy", interp="log"):
        edges = np.asarray(to_value(energy_edges, "energy"), dtype=float)
        if unit is not None:
            edges = edges * to_value(f"1 {unit}", "energy")
        if np.any(edges <= 0) or np.any(np.diff(edges) <= 0):
            raise ValueError("MapAxis: energy edges must be positive and increasing")
        return cls(edges, interp=interp, name=name, node_type="edges", unit="TeV")

    @classmethod
    def from_edges(cls, edges, name="", unit="", interp="lin"):
        return cls(edges, interp=interp, name=name, node_type="edges", unit=unit)

    @classmethod
    def from_nodes(cls, nodes, name="", unit="", interp="lin"):
        """maps/axes.py:663-685: bin centres given, edges half-way between them in the axis' interpolation scale."""
        if len(nodes) < 1:
            raise ValueError("Nodes array must have at least one element.")
        return cls(nodes, interp=interp, name=name, node_type="center", unit=unit)

    @classmethod
    def from_bounds(cls, lo_bnd, hi_bnd, nbin, interp="lin", node_type="edges", name="", unit=""):
        nnode = int(nbin) + 1 if node_type == "edges" else int(nbin)
        if interp == "lin":
            nodes = np.linspace(lo_bnd, hi_bnd, nnode)
        elif interp == "log":
            nodes = np.geomspace(lo_bnd, hi_bnd, nnode)
        elif interp == "sqrt":
            nodes = np.linspace(lo_bnd**0.5, hi_bnd**0.5, nnode) ** 2.0
        else:
            raise ValueError(f"Invalid interp: {interp}")
        return cls(nodes, interp=interp, name=name, node_type=node_type, unit=unit)

    # -- pix <-> coord ------------------------------------------------------------------------
    def _scale(self, v):
        if self.interp == "log":
            return np.log(np.clip(v, np.finfo(float).tiny, np.inf))
        if self.interp == "sqrt":
            return np.sqrt(v)
        return np.asarray(v, dtype=float)

    def _inverse(self, v):
        if self.interp == "log":
            return np.exp(v)
        if self.interp == "sqrt":
            return np.power(v, 2)
        return v

    def pix_to_coord(self, pix):
        """maps/axes.py:798-813: linear interpolation of the scaled nodes over pixel index (extrapolating)."""
        pix = np.asarray(pix, dtype=float) - self._pix_offset
        x = np.arange(len(self._nodes), dtype=float)
        y = self._scale(self._nodes)
        if len(x) == 1:
            return self._inverse(np.full_like(pix, y[0]))
        idx = np.clip(np.floor(pix).astype(int), 0, len(x) - 2)
        val = y[idx] + (pix - x[idx]) * (y[idx + 1] - y[idx]) / (x[idx + 1] - x[idx])
        return self._inverse(val)

    def coord_to_pix(self, coord):
        y = self._scale(self._nodes)
        c = self._scale(np.asarray(coord, dtype=float))
        return np.interp(c, y, np.arange(len(y), dtype=float)) + self._pix_offset

    def coord_to_idx(self, coord, clip=False):
        pix = self.coord_to_pix(coord)
        idx = np.round(pix).astype(int) if self.node_type == "center" else np.floor(pix + 0.5).astype(int)
        c = np.asarray(coord, dtype=float)
        outside = (c < self.edges[0]) | (c >= self.edges[-1])
        if clip:
            return np.clip(idx, 0, self.nbin - 1)
        return np.where(outside, -1, idx)

    @property
    def nbin(self):
        return self._nbin

    @property
    def edges(self):
        return self.pix_to_coord(np.arange(self.nbin + 1, dtype=float) - 0.5)

    @property
    def edges_min(self):
        return self.edges[:-1]

    @property
    def edges_max(self):
        return self.edges[1:]

    @property
    def center(self):
        return self.pix_to_coord(np.arange(self.nbin, dtype=float))

    @property
    def bin_width(self):
        return np.diff(self.edges)

    @property
    def bounds(self):
        e = self.edges
        return e[0], e[-1]

    def copy(self, **kwargs):
        new = _copy.deepcopy(self)
        for k, v in kwargs.items():
            setattr(new, k, v)
        return new

    def upsample(self, factor):
        """maps/axes.py:1087-1117: factor * nbin bins between the same bounds (edges axis), (nbin - 1) * factor + 1 nodes
        between the same end nodes (centre-node axis); same interpolation."""
        if self.node_type != "edges":
            pix = self.coord_to_pix(self.center)
            pix_new = np.linspace(pix.min(), pix.max(), int((self.nbin - 1) * factor) + 1)
            return MapAxis.from_nodes(self.pix_to_coord(pix_new), name=self.name, unit=self.unit, interp=self.interp)
        edges = self.edges
        return MapAxis.from_bounds(edges[0], edges[-1], self.nbin * factor, interp=self.interp, name=self.name,
                                   unit=self.unit)

    def squash(self):
        e = self.edges
        return MapAxis([e[0], e[-1]], interp=self.interp, name=self.name, unit=self.unit)

    def __eq__(self, other):
        return (isinstance(other, MapAxis) and self.name == other.name and self.nbin == other.nbin
                and self.node_type == other.node_type and np.allclose(self._nodes, other._nodes, rtol=1e-6, atol=1e-6))

    def __repr__(self):
        lo, hi = self.bounds
        return f"MapAxis(name={self.name!r}, nbin={self.nbin}, bounds=({lo:.4g}, {hi:.4g}) {self.unit}, interp={self.interp!r})"


class MapAxes(list):
    @property
    def names(self):
        return [ax.name for ax in self]

    def __getitem__(self, key):
        if isinstance(key, str):
            for ax in self:
                if ax.name == key:
                    return ax
            raise KeyError(f"No axis named {key!r}; available: {self.names}")
        return list.__getitem__(self, key)


def _pair(value, cast=float):
    if isinstance(value, (tuple, list, np.ndarray)):
        return cast(to_value(value[0], "angle") if cast is float else value[0]), \
            cast(to_value(value[1], "angle") if cast is float else value[1])
    v = cast(to_value(value, "angle") if cast is float else value)
    return v, v


class NoOverlapError(ValueError):
    pass


def round_up_to_odd(f):
    return (np.ceil(f) // 2 * 2 + 1).astype(int)


class WcsGeom:
    """WCS map geometry (maps/wcs/geom.py:294-417), square pixels, CRVAL = ``skydir`` as ``WcsGeom.create`` sets it
    (geom.py:370-410): CAR (plate carree; separable about the equator, oblique about any other reference point: the
    native origin is tilted up to the reference latitude, which is what the default LONPOLE / LATPOLE of FITS WCS paper
    II, Calabretta & Greisen 2002, section 2.4 and eqs. 8-10 give) or TAN (gnomonic, eqs. 5, 54, 55 with the native pole
    at the reference point and LONPOLE = 180 deg)."""

    def __init__(self, npix, binsz, crval, crpix, frame="icrs", axes=None, proj="CAR"):
        self._npix = (int(npix[0]), int(npix[1]))
        self._binsz = float(binsz)
        self._crval = (float(crval[0]), float(crval[1]))
        self._crpix = (float(crpix[0]), float(crpix[1]))
        self.frame = frame
        self.projection = proj
        self.axes = MapAxes(axes or [])
        if proj not in ("CAR", "TAN"):
            raise NotImplementedError(f"gammapy_b200 implements the CAR and TAN projections only, got {proj!r}")
        if proj == "CAR" and not -90.0 < self._crval[1] < 90.0:
            raise ValueError("CAR: reference latitude must lie inside (-90, 90) deg")

    @classmethod
    def create(cls, npix=None, binsz=0.5, proj="CAR", frame="icrs", refpix=None, axes=None, skydir=None, width=None):
        crval = as_lonlat(skydir)
        if isinstance(skydir, SkyCoord):
            frame = skydir.frame
        bx, by = _pair(binsz)
        if bx != by:
            raise NotImplementedError("gammapy_b200 supports square pixels only")
        if npix is None and width is None:
            width = (360.0, 180.0)
        if npix is None:
            wx, wy = _pair(width)
            npix = (int(np.rint(wx / bx)), int(np.rint(wy / by)))
        else:
            npix = _pair(npix, int)
        if refpix is None:
            refpix = ((npix[0] + 1) / 2.0, (npix[1] + 1) / 2.0)
        return cls(npix, bx, crval, refpix, frame=frame, axes=axes, proj=proj)

    # -- basic properties ------------------------------------------------------------------------
    @property
    def npix(self):
        return self._npix

    @property
    def binsz(self):
        return self._binsz

    @property
    def crval(self):
        return self._crval

    @property
    def crpix(self):
        return self._crpix

    @property
    def is_image(self):
        return len(self.axes) == 0

    @property
    def data_shape_image(self):
        return (self._npix[1], self._npix[0])

    @property
    def data_shape(self):
        return tuple(ax.nbin for ax in reversed(self.axes)) + self.data_shape_image

    @property
    def width(self):
        return (self._binsz * self._npix[0], self._binsz * self._npix[1])

    @property
    def pixel_scales(self):
        return np.array([self._binsz, self._binsz])

    @property
    def center_pix(self):
        return ((self._npix[0] - 1) / 2.0, (self._npix[1] - 1) / 2.0)

    @property
    def center_skydir(self):
        lon, lat = self.pix_to_coord(self.center_pix)[:2]
        return SkyCoord(float(lon), float(lat), frame=self.frame)

    def _init_copy(self, **kw):
        d = dict(npix=self._npix, binsz=self._binsz, crval=self._crval, crpix=self._crpix, frame=self.frame,
                 axes=_copy.deepcopy(list(self.axes)), proj=self.projection)
        d.update(kw)
        return WcsGeom(**d)

    def copy(self, **kw):
        return self._init_copy(**kw)

    def to_image(self):
        return self._init_copy(axes=[])

    def to_cube(self, axes):
        return self._init_copy(axes=list(self.axes) + list(axes))

    def drop(self, axis_name):
        return self._init_copy(axes=[ax for ax in self.axes if ax.name != axis_name])

    def __eq__(self, other):
        return (isinstance(other, WcsGeom) and self._npix == other._npix and self._binsz == other._binsz
                and self.projection == other.projection and np.allclose(self._crval, other._crval) and np.allclose(self._crpix, other._crpix)
                and len(self.axes) == len(other.axes) and all(a == b for a, b in zip(self.axes, other.axes)))

    def is_aligned(self, other):
        if self._binsz != other._binsz or self.projection != other.projection or not np.allclose(self._crval, other._crval):
            return False
        d = np.array(self._crpix) - np.array(other._crpix)
        return bool(np.allclose(d, np.round(d)))

    # -- coordinates ---------------------------------------------------------------------------------
    def pix_to_coord(self, pix):
        """wcs_pix2world(px, py, 0) (+ non-spatial axes)."""
        px = np.asarray(pix[0], dtype=float)
        py = np.asarray(pix[1], dtype=float)
        if self.projection == "TAN":
            x = (-self._binsz) * (px + 1.0 - self._crpix[0])
            y = self._binsz * (py + 1.0 - self._crpix[1])
            x, y = np.broadcast_arrays(x, y)
            r = np.hypot(x, y)
            phi = np.arctan2(x, -y)
            theta = np.arctan2(180.0 / np.pi, r)
            dphi = phi - np.pi  # LONPOLE = 180 deg
            lat_p = self._crval[1] * DEG
            st, ct = np.sin(theta), np.cos(theta)
            lat = np.arcsin(np.clip(st * np.sin(lat_p) + ct * np.cos(lat_p) * np.cos(dphi), -1.0, 1.0))
            lon = self._crval[0] + np.arctan2(-ct * np.sin(dphi), st * np.cos(lat_p) - ct * np.sin(lat_p) * np.cos(dphi)) / DEG
            lon = np.where(r == 0.0, self._crval[0], lon)
            lat = np.where(r == 0.0, self._crval[1], lat / DEG)
        else:
            lon = self._crval[0] + (-self._binsz) * (px + 1.0 - self._crpix[0])
            lat = self._binsz * (py + 1.0 - self._crpix[1])
            if self._crval[1] != 0.0:  # oblique: (lon - crval, lat) are the native (phi, theta); tilt by the reference latitude
                lon, lat = np.broadcast_arrays(lon, lat)
                phi, theta, lat0 = (lon - self._crval[0]) * DEG, lat * DEG, self._crval[1] * DEG
                ct = np.cos(theta)
                xs = ct * np.cos(phi) * np.cos(lat0) - np.sin(theta) * np.sin(lat0)
                ys = ct * np.sin(phi)
                zs = ct * np.cos(phi) * np.sin(lat0) + np.sin(theta) * np.cos(lat0)
                lon = self._crval[0] + np.arctan2(ys, xs) / DEG
                lat = np.arctan2(zs, np.hypot(xs, ys)) / DEG
        out = [lon, lat]
        for ax, p in zip(self.axes, pix[2:]):
            out.append(ax.pix_to_coord(p))
        return tuple(out)

    def coord_to_pix(self, coords):
        if isinstance(coords, SkyCoord):
            coords = coords.transform_to(self.frame)  # (wcs_utils.skycoord_to_pixel brings the position into the map frame)
            coords = (coords.lon, coords.lat)
        lon = np.asarray(coords[0], dtype=float)
        lat = np.asarray(coords[1], dtype=float)
        if self.projection == "TAN":
            dl = (lon - self._crval[0]) * DEG
            la, lat_p = lat * DEG, self._crval[1] * DEG
            # gnomonic: x = cos(lat) sin(dlon) / s, y = (sin(lat) cos(lat_p) - cos(lat) sin(lat_p) cos(dlon)) / s with
            # s = sin(theta) = cos of the angular distance to the reference point (s <= 0: behind the tangent plane)
            s_ = np.sin(la) * np.sin(lat_p) + np.cos(la) * np.cos(lat_p) * np.cos(dl)
            with np.errstate(divide="ignore", invalid="ignore"):
                x = np.where(s_ > 0, (np.cos(la) * np.sin(dl) / s_) / DEG, np.nan)
                y = np.where(s_ > 0, ((np.sin(la) * np.cos(lat_p) - np.cos(la) * np.sin(lat_p) * np.cos(dl)) / s_) / DEG, np.nan)
            px = x / (-self._binsz) + self._crpix[0] - 1.0
            py = y / self._binsz + self._crpix[1] - 1.0
        elif self._crval[1] != 0.0:  # oblique CAR: tilt back by the reference latitude, pixel = native (phi, theta) / cdelt
            dl, la, lat0 = (lon - self._crval[0]) * DEG, lat * DEG, self._crval[1] * DEG
            xs, ys, zs = np.cos(la) * np.cos(dl), np.cos(la) * np.sin(dl), np.sin(la)
            xn = xs * np.cos(lat0) + zs * np.sin(lat0)
            zn = -xs * np.sin(lat0) + zs * np.cos(lat0)
            px = (np.arctan2(ys, xn) / DEG) / (-self._binsz) + self._crpix[0] - 1.0
            py = (np.arctan2(zn, np.hypot(xn, ys)) / DEG) / self._binsz + self._crpix[1] - 1.0
        else:
            dlon = (lon - self._crval[0] + 180.0) % 360.0 - 180.0
            px = dlon / (-self._binsz) + self._crpix[0] - 1.0
            py = lat / self._binsz + self._crpix[1] - 1.0
        out = [px, py]
        for ax, c in zip(self.axes, coords[2:]):
            out.append(ax.coord_to_pix(c))
        return tuple(out)

    def get_coord(self, mode="center", sparse=False):
        """Dict of broadcastable coordinate arrays: lon, lat [deg] and one entry per non-spatial axis."""
        nx, ny = self._npix
        if mode == "edges":
            x, y = np.arange(-0.5, nx, dtype=float), np.arange(-0.5, ny, dtype=float)
        else:
            x, y = np.arange(nx, dtype=float), np.arange(ny, dtype=float)
        nd = len(self.axes) + 2
        if self.projection == "TAN" or self._crval[1] != 0.0:  # not separable: full (ny, nx) images of both coordinates
            lon, lat = self.pix_to_coord((x[np.newaxis, :], y[:, np.newaxis]))[:2]
            shape2 = [1] * (nd - 2) + [len(y), len(x)]
            coords = {"lon": lon.reshape(shape2), "lat": lat.reshape(shape2)}
            for i, ax in enumerate(self.axes):
                shape = [1] * nd
                shape[nd - 3 - i] = -1
                coords[ax.name] = ax.center.reshape(shape)
            if not sparse:
                full = self.data_shape if mode != "edges" else tuple(ax.nbin for ax in reversed(self.axes)) + (len(y), len(x))
                coords = {k: np.broadcast_to(v, full) if k in ("lon", "lat") else v for k, v in coords.items()}
            return coords
        lon, lat = self.pix_to_coord((x, y))[:2]
        shape_x = [1] * nd
        shape_x[-1] = -1
        shape_y = [1] * nd
        shape_y[-2] = -1
        coords = {"lon": lon.reshape(shape_x), "lat": lat.reshape(shape_y)}
        for i, ax in enumerate(self.axes):
            shape = [1] * nd
            shape[nd - 3 - i] = -1
            coords[ax.name] = ax.center.reshape(shape)
        if not sparse:
            full = self.data_shape if mode != "edges" else tuple(ax.nbin for ax in reversed(self.axes)) + (len(y), len(x))
            coords = {k: np.broadcast_to(v, full) if k in ("lon", "lat") else v for k, v in coords.items()}
        return coords

    def separation(self, center):
        """Angular separation [deg] of every pixel centre from ``center`` (maps/wcs/geom.py:868-884)."""
        lon0, lat0 = as_lonlat(center.transform_to(self.frame) if isinstance(center, SkyCoord) else center)
        c = self.to_image().get_coord()
        return angular_separation(c["lon"] * DEG, c["lat"] * DEG, lon0 * DEG, lat0 * DEG) / DEG

    def contains(self, coords):
        px, py = self.coord_to_pix(coords)[:2]
        return (px >= -0.5) & (px < self._npix[0] - 0.5) & (py >= -0.5) & (py < self._npix[1] - 0.5)

    def solid_angle(self):
        """maps/wcs/geom.py:816-854 (planar triangles on the pixel corners), sr, shape (ny, nx)."""
        c = self.to_image().get_coord(mode="edges")
        lon, lat = c["lon"] * DEG, c["lat"] * DEG
        lo, hi = slice(None, -1), slice(1, None)

        def pt(sy, sx):
            return lon[sy, sx], lat[sy, sx]

        low_left, low_right, up_left, up_right = pt(lo, lo), pt(hi, lo), pt(lo, hi), pt(hi, hi)
        low = angular_separation(*low_left, *low_right)
        left = angular_separation(*low_left, *up_left)
        up = angular_separation(*up_left, *up_right)
        right = angular_separation(*low_right, *up_right)
        angle_low_right = position_angle(*low_right, *up_right) - position_angle(*low_right, *low_left)
        angle_up_left = position_angle(*up_left, *up_right) - position_angle(*low_left, *up_left)
        area = 0.5 * low * right * np.sin(angle_low_right) + 0.5 * up * left * np.sin(angle_up_left)
        return np.abs(area)

    def bin_volume(self):
        vol = self.solid_angle()
        for ax in self.axes:
            vol = ax.bin_width.reshape((-1,) + (1,) * vol.ndim) * vol
        return vol

    # -- derived geometries ------------------------------------------------------------------------------
    def upsample(self, factor):
        return self._init_copy(
            npix=(self._npix[0] * factor, self._npix[1] * factor), binsz=self._binsz / factor,
            crpix=((self._crpix[0] - 0.5) * factor + 0.5, (self._crpix[1] - 0.5) * factor + 0.5))

    def downsample(self, factor):
        if self._npix[0] % factor or self._npix[1] % factor:
            raise ValueError(f"Spatial shape not divisible by factor {factor!r} in all axes.")
        return self._init_copy(
            npix=(self._npix[0] // factor, self._npix[1] // factor), binsz=self._binsz * factor,
            crpix=((self._crpix[0] - 0.5) / factor + 0.5, (self._crpix[1] - 0.5) / factor + 0.5))

    def to_odd_npix(self, max_radius=None):
        """maps/wcs/geom.py:1106-1138."""
        width = max(self.width) if max_radius is None else 2 * to_value(max_radius, "angle")
        npix = int(round_up_to_odd(np.float64(width / self._binsz)))
        c = self.center_skydir
        return WcsGeom.create(skydir=(c.lon, c.lat), binsz=self._binsz, npix=npix, proj=self.projection,
                              frame=self.frame, axes=list(self.axes))

    def cutout_slices_for(self, position, width, mode="trim", odd_npix=False, min_npix=1):
        """Parent (y, x) slices of ``WcsGeom.cutout`` (maps/wcs/geom.py:886-930; astropy Cutout2D rule)."""
        wx, wy = _pair(width)
        width_npix = np.clip(np.array([wx, wy]) / self._binsz, min_npix, None)
        if odd_npix:
            width_npix = round_up_to_odd(width_npix)
        shape = (int(np.round(width_npix[1])), int(np.round(width_npix[0])))
        lon, lat = as_lonlat(position.transform_to(self.frame) if isinstance(position, SkyCoord) else position)
        px, py = self.coord_to_pix((lon, lat))[:2]
        pos = (float(py), float(px))
        if not all(np.isfinite(pos)):
            raise ValueError("Input position contains invalid values (NaNs or infs).")
        large = self.data_shape_image
        emin = [int(np.ceil(p - s / 2.0)) for p, s in zip(pos, shape)]
        emax = [int(np.ceil(p + s / 2.0)) for p, s in zip(pos, shape)]
        if any(e <= 0 for e in emax) or any(e >= l for e, l in zip(emin, large)):
            raise NoOverlapError("Arrays do not overlap.")
        return tuple(slice(max(0, a), min(l, b)) for a, b, l in zip(emin, emax, large))

    def cutout(self, position, width, mode="trim", odd_npix=False, min_npix=1):
        ys, xs = self.cutout_slices_for(position, width, mode=mode, odd_npix=odd_npix, min_npix=min_npix)
        return self._init_copy(npix=(xs.stop - xs.start, ys.stop - ys.start),
                               crpix=(self._crpix[0] - xs.start, self._crpix[1] - ys.start))

    def __repr__(self):
        c = self.center_skydir
        return (f"WcsGeom(axes={['lon', 'lat'] + self.axes.names}, shape={self.data_shape[::-1]}, frame={self.frame!r}, "
                f"proj={self.projection!r}, center=({c.lon:.3f}, {c.lat:.3f}) deg, binsz={self._binsz} deg)")


class Map:
    """N-dim sky map = geometry + numpy data (maps/core.py, maps/wcs/ndmap.py)."""

    def __init__(self, geom, data=None, unit="", dtype="float32", meta=None):
        self.geom = geom
        self.unit = unit
        self.meta = meta or {}
        if data is None:
            data = np.zeros(geom.data_shape, dtype=dtype)
        else:
            data = np.asarray(data)
            if data.shape != geom.data_shape:
                data = np.broadcast_to(data, geom.data_shape).copy()
        self.data = data

    @staticmethod
    def from_geom(geom, meta=None, data=None, unit="", dtype="float32"):
        """maps/core.py:269-305: default dtype float32."""
        return WcsNDMap(geom, data=data, unit=unit, dtype=dtype, meta=meta)

    @staticmethod
    def create(**kwargs):
        unit = kwargs.pop("unit", "")
        dtype = kwargs.pop("dtype", "float32")
        data = kwargs.pop("data", None)
        return WcsNDMap(WcsGeom.create(**kwargs), data=data, unit=unit, dtype=dtype)

    @property
    def quantity(self):
        return self.data

    @quantity.setter
    def quantity(self, val):
        self.data = np.asarray(val)

    @property
    def is_mask(self):
        return self.data.dtype == bool

    def copy(self, **kw):
        new = _copy.copy(self)
        new.data = self.data.copy()
        for k, v in kw.items():
            setattr(new, k, v)
        return new

    def _arith(self, op, other, inplace=False):
        other_data = other.data if isinstance(other, Map) else np.asarray(other)
        if isinstance(other, Map) and other.geom.data_shape != self.geom.data_shape:
            raise ValueError("Map Arithmetic: Inconsistent geometries.")
        out = self if inplace else _copy.copy(self)
        out.data = op(self.data, other_data)
        return out

    def __add__(self, o):
        return self._arith(np.add, o)

    def __iadd__(self, o):
        return self._arith(np.add, o, True)

    def __sub__(self, o):
        return self._arith(np.subtract, o)

    def __mul__(self, o):
        return self._arith(np.multiply, o)

    def __imul__(self, o):
        return self._arith(np.multiply, o, True)

    def __truediv__(self, o):
        return self._arith(np.true_divide, o)

    def __and__(self, o):
        return self._arith(np.logical_and, o)

    def __or__(self, o):
        return self._arith(np.logical_or, o)

    def __invert__(self):
        out = _copy.copy(self)
        out.data = np.logical_not(self.data)
        return out

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)

    def sum_over_axes(self, keepdims=False):
        return self.reduce_over_axes(np.add, keepdims=keepdims)

    def reduce_over_axes(self, func=np.add, keepdims=False):
        data = self.data
        nax = len(self.geom.axes)
        for _ in range(nax):
            data = func.reduce(data, axis=0)
        geom = self.geom.to_image()
        return WcsNDMap(geom, data=data, unit=self.unit)

    def cutout(self, position, width, mode="trim", odd_npix=False, min_npix=1):
        ys, xs = self.geom.cutout_slices_for(position, width, mode=mode, odd_npix=odd_npix, min_npix=min_npix)
        geom = self.geom.cutout(position, width, mode=mode, odd_npix=odd_npix, min_npix=min_npix)
        return WcsNDMap(geom, data=self.data[..., ys, xs].copy(), unit=self.unit)

    def get_by_coord(self, coords):
        pix = self.geom.coord_to_pix(coords)
        idx = tuple(np.round(p).astype(int) for p in reversed(pix))
        return self.data[idx]

    def __repr__(self):
        return f"{type(self).__name__}(geom={self.geom!r}, unit={self.unit!r}, dtype={self.data.dtype})"


class WcsNDMap(Map):
    def convolve(self, kernel, method="direct", mode="same"):
        """``WcsNDMap.convolve`` (maps/wcs/ndmap.py:920-1010) of an image with a PSF kernel cube, on the GPU
        (direct fp32 stencil instead of the reference's fp32 FFT)."""
        from ._lib import Context, check
        import ctypes as C

        if mode != "same":
            raise NotImplementedError("WcsNDMap.convolve: only mode='same' is implemented")
        kdata = kernel.data if hasattr(kernel, "data") and not isinstance(kernel, np.ndarray) else kernel
        kdata = np.ascontiguousarray(np.asarray(kdata), dtype=np.float32)
        if kdata.ndim == 2:
            kdata = kdata[np.newaxis]
        if not self.geom.is_image:
            raise NotImplementedError("WcsNDMap.convolve: image input only")
        img = np.ascontiguousarray(self.data, dtype=np.float32)
        out = np.empty((kdata.shape[0],) + img.shape, dtype=np.float32)
        ctx = Context.get()
        check(ctx.lib.gpy_psf_convolve(ctx.handle, img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1],
                                       kdata.ctypes.data_as(C.c_void_p), kdata.shape[0], kdata.shape[-1],
                                       out.ctypes.data_as(C.c_void_p)))
        axes = list(getattr(getattr(kernel, "psf_kernel_map", None), "geom", self.geom).axes) if kdata.shape[0] > 1 else []
        if axes:
            return WcsNDMap(self.geom.to_cube(axes), data=out, unit=self.unit)
        return WcsNDMap(self.geom, data=out[0], unit=self.unit)
