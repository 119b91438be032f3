"""Static IRF producers: the PSF kernel cube ``(nT, K, K)`` and the edisp matrix ``(nT, nR)``.

These are *inputs* of the likelihood hot path, built once on the host with numpy.
Reference: irf/psf/map.py (PSFMap.from_gauss :397-448, get_psf_kernel :249-337, _psf_upsampling_factor
:23-37), irf/psf/kernel.py (PSFKernel :12-126), irf/psf/core.py (containment_radius :39-87),
irf/edisp/map.py (EDispKernelMap, get_overlap_fraction :11-20), irf/edisp/kernel.py (EDispKernel).

A PSFMap / EDispKernelMap is position independent (one radial profile per true energy / one matrix, as ``from_gauss``
makes them) or carries a spatial grid (``geom``: a CAR image geometry, data with two trailing (lat, lon) axes); the
evaluators then look their IRFs up at the source position -- on the device, ``gpy_dataset_set_irf_maps`` -- and the
methods here are the host mirror of that lookup (bilinear profile, nearest-pixel matrix).
``EDispKernel.from_gauss`` uses the closed-form Gaussian migration integral per (true, reco) bin rather
than the reference's 200-bin migra interpolation (irf/edisp/core.py:88-136 + makers/utils.py:345-388);
its output is the static matrix both this package and the oracle consume.
"""
import numpy as np
import scipy.special

from .maps import Map, MapAxis, WcsGeom
from .utils import DEG, angular_separation, to_value

__all__ = ["PSFMap", "PSFKernel", "EDispKernelMap", "EDispKernel", "RAD_AXIS_DEFAULT"]

RAD_MAX = 0.66
RAD_AXIS_DEFAULT = MapAxis.from_bounds(0, RAD_MAX, nbin=66, node_type="edges", name="rad", unit="deg")
PSF_MAX_OVERSAMPLING = 4  # irf/psf/map.py:20


class PSFKernel:
    """PSF kernel cube held in a Map (irf/psf/kernel.py:12-126); planes normalised to 1."""

    def __init__(self, psf_kernel_map, normalize=True):
        self._psf_kernel_map = psf_kernel_map
        if normalize:
            self.normalize()

    def normalize(self):
        """irf/psf/kernel.py:70-80."""
        data = self.psf_kernel_map.data
        axis = (0, 1) if self.psf_kernel_map.geom.is_image else (1, 2)
        with np.errstate(divide="ignore", invalid="ignore"):
            data = np.nan_to_num(data / data.sum(axis=axis, keepdims=True))
        self.psf_kernel_map.data = data

    @property
    def data(self):
        return self._psf_kernel_map.data

    @property
    def psf_kernel_map(self):
        return self._psf_kernel_map

    @classmethod
    def from_gauss(cls, geom, sigma, max_radius=None, factor=4):
        """Gaussian kernel integrated on a ``factor``-oversampled grid (irf/psf/kernel.py:99-160)."""
        sigma = to_value(sigma, "angle")
        if max_radius is None:
            max_radius = 5 * np.max(sigma)
        g = geom.to_image().to_odd_npix(max_radius=max_radius)
        up = g.upsample(factor)
        c = up.get_coord()
        center = g.center_skydir
        rad = angular_separation(c["lon"] * DEG, c["lat"] * DEG, center.lon * DEG, center.lat * DEG)
        sig = np.atleast_1d(sigma) * DEG
        planes = []
        for s in sig:
            val = np.exp(-0.5 * (rad / s) ** 2)
            ny, nx = val.shape
            planes.append(val.reshape(ny // factor, factor, nx // factor, factor).sum(axis=(1, 3)))
        data = np.array(planes)
        axes = list(geom.axes) if len(sig) > 1 else []
        if not axes:
            data = data[0]
        return cls(Map.from_geom(g.to_cube(axes) if axes else g, data=data, dtype=float))


def _irf_pixel(geom, lon, lat):
    """Fractional pixel of (lon, lat) [deg] in the IRF image grid, clamped to the grid
    (``PSFMap._get_nearest_valid_position`` keeps the lookup inside the map)."""
    dlon = (lon - geom.crval[0] + 180.0) % 360.0 - 180.0
    px = dlon / (-geom.binsz) + geom.crpix[0] - 1.0
    py = lat / geom.binsz + geom.crpix[1] - 1.0
    return float(np.clip(px, 0.0, geom.npix[0] - 1.0)), float(np.clip(py, 0.0, geom.npix[1] - 1.0))


class PSFMap:
    """PSF radial profile ``psf(energy_true, rad)`` in sr-1 (irf/psf/map.py): one for the whole field
    (``data`` (nT, nrad)) or one per pixel of the image geometry ``geom`` (``data`` (nT, nrad, ny, nx))."""

    energy_name = "energy_true"

    def __init__(self, energy_axis_true, rad_axis, data, geom=None):
        self.energy_axis_true = energy_axis_true
        self.rad_axis = rad_axis
        self.data = np.asarray(data, dtype=float)  # at rad_axis.center
        self.geom = None if geom is None else geom.to_image()
        if self.data.ndim == 4:
            if self.geom is None:
                raise ValueError("a PSF map with spatial axes needs its image geometry")
            if self.data.shape[2:] != (self.geom.npix[1], self.geom.npix[0]):
                raise ValueError("PSF map data do not match the image geometry")
        elif self.data.ndim != 2:
            raise ValueError("PSF map data must be (n_energy_true, n_rad[, ny, nx])")
        if self.data.shape[:2] != (energy_axis_true.nbin, rad_axis.nbin):
            raise ValueError("PSF map data do not match the energy / rad axes")
        self.has_single_spatial_bin = self.data.ndim == 2

    def at_position(self, position=None):
        """Position-independent PSFMap with the profile at ``position`` ((lon, lat) [deg]; default: the grid centre):
        bilinear in the IRF pixel grid (``Map.interp_by_coord(method="linear")``, irf/psf/map.py:154-170, 310-320)."""
        if self.has_single_spatial_bin:
            return self
        if position is None:
            c = self.geom.center_skydir
            position = (c.lon, c.lat)
        lon, lat = (position.lon, position.lat) if hasattr(position, "lon") else position
        px, py = _irf_pixel(self.geom, float(lon), float(lat))
        ny, nx = self.data.shape[2:]
        x0 = int(min(np.floor(px), max(nx - 2, 0)))
        y0 = int(min(np.floor(py), max(ny - 2, 0)))
        x1, y1 = min(x0 + 1, nx - 1), min(y0 + 1, ny - 1)
        tx, ty = px - x0, py - y0
        d = self.data
        prof = ((1 - ty) * ((1 - tx) * d[..., y0, x0] + tx * d[..., y0, x1])
                + ty * ((1 - tx) * d[..., y1, x0] + tx * d[..., y1, x1]))
        return PSFMap(self.energy_axis_true, self.rad_axis, prof)

    @classmethod
    def from_gauss(cls, energy_axis_true, rad_axis=None, sigma=0.1, geom=None):
        """irf/psf/map.py:397-448: Gauss2DPDF (utils/gauss.py) evaluated at the rad bin centres."""
        if rad_axis is None:
            rad_axis = RAD_AXIS_DEFAULT.copy()
        sigma = np.atleast_1d(to_value(sigma, "angle")).astype(float)
        if len(sigma) not in (1, energy_axis_true.nbin):
            raise ValueError("sigma must be a scalar or one value per true energy bin")
        sigma = np.broadcast_to(sigma, (energy_axis_true.nbin,)) * DEG
        rad = rad_axis.center * DEG
        s2 = (sigma * sigma)[:, np.newaxis]
        data = 1 / (2 * np.pi * s2) * np.exp(-0.5 * (rad[np.newaxis, :] ** 2) / s2)
        return cls(energy_axis_true, rad_axis, data)

    def _interp_rad(self, rad_deg):
        """Linear interpolation in rad of each energy row, extrapolating with the edge slope
        (Map.interp_by_coord(method='linear', fill_value=None) on the rad nodes)."""
        x = self.rad_axis.center
        r = np.asarray(rad_deg, dtype=float)
        idx = np.clip(np.searchsorted(x, r) - 1, 0, len(x) - 2)
        t = (r - x[idx]) / (x[idx + 1] - x[idx])
        return self.data[:, idx] * (1 - t) + self.data[:, idx + 1] * t

    def containment(self, rad, energy_idx=None):
        """Fraction within ``rad`` [deg]: cumulative integral of psf * 2 pi rad over the rad bins."""
        edges = self.rad_axis.edges * DEG
        centers = self.rad_axis.center * DEG
        ring = np.pi * (edges[1:] ** 2 - edges[:-1] ** 2)
        cum = np.concatenate([np.zeros((self.data.shape[0], 1)), np.cumsum(self.data * ring, axis=1)], axis=1)
        del centers
        r = np.asarray(rad, dtype=float) * DEG
        out = np.array([np.interp(r, edges, row) for row in cum])
        return out if energy_idx is None else out[energy_idx]

    def containment_radius(self, fraction, factor=20):
        """irf/psf/core.py:39-87: grid search on the ``factor``-upsampled rad axis, one radius per energy."""
        rad = self.rad_axis.upsample(factor).center
        cont = self.containment(rad)  # (nT, nrad_up)
        idx = np.argmin(np.abs(cont - fraction), axis=1)
        return rad[idx]

    def get_psf_kernel(self, geom, position=None, max_radius=None, containment=0.999, precision_factor=12):
        """irf/psf/map.py:249-337 (host mirror; the evaluators of a dataset do this lookup on the device)."""
        if not self.has_single_spatial_bin:
            return self.at_position(position).get_psf_kernel(geom, None, max_radius, containment, precision_factor)
        geom = geom.to_image()
        radii = self.containment_radius(containment)
        if max_radius is None:
            max_radius = np.max(radii)
        else:
            max_radius = to_value(max_radius, "angle")
            radii = np.minimum(radii, max_radius)
        r68 = self.containment_radius(0.68)
        pix = np.max(geom.pixel_scales)
        kgeom = geom.to_odd_npix(max_radius=max_radius)
        K = kgeom.npix[0]
        nT = self.energy_axis_true.nbin
        data = np.zeros((nT, K, K), dtype=float)
        center = kgeom.center_skydir
        for t in range(nT):
            base_factor = 2 * r68[t] / pix
            with np.errstate(divide="ignore"):
                factor = int(np.minimum(int(np.ceil(precision_factor / base_factor)) if base_factor > 0 else
                                        PSF_MAX_OVERSAMPLING, PSF_MAX_OVERSAMPLING))
            cut = kgeom.to_odd_npix(max_radius=radii[t])
            up = cut.upsample(factor)
            c = up.get_coord()
            rad = angular_separation(c["lon"] * DEG, c["lat"] * DEG, center.lon * DEG, center.lat * DEG) / DEG
            vals = np.clip(self._interp_rad(rad.ravel())[t].reshape(rad.shape), 0, np.inf)
            # Map.from_geom(data=psf value) then downsample(preserve_counts=True): block sum
            n = cut.npix[0]
            img = vals.reshape(n, factor, n, factor).sum(axis=(1, 3))
            off = (K - n) // 2
            data[t, off:off + n, off:off + n] = img
        kmap = Map.from_geom(kgeom.to_cube([self.energy_axis_true]), data=data, dtype=float)
        return PSFKernel(kmap, normalize=True)


def get_overlap_fraction(energy_axis, energy_axis_true):
    """irf/edisp/map.py:11-20."""
    a_min = energy_axis.edges[:-1]
    a_max = energy_axis.edges[1:]
    b_min = energy_axis_true.edges[:-1][:, np.newaxis]
    b_max = energy_axis_true.edges[1:][:, np.newaxis]
    xmin = np.fmin(a_max, b_max)
    xmax = np.fmax(a_min, b_min)
    return np.clip(xmin - xmax, 0, np.inf) / (b_max - b_min)


class EDispKernel:
    """Energy dispersion matrix, rows = true energy, columns = reco energy (irf/edisp/kernel.py:71-77)."""

    def __init__(self, axes, data):
        self.axes = {ax.name: ax for ax in axes}
        self.data = np.asarray(data, dtype=float)

    @property
    def pdf_matrix(self):
        return self.data

    @classmethod
    def from_diagonal_response(cls, energy_axis_true, energy_axis=None):
        """irf/edisp/kernel.py:163-211."""
        if energy_axis is None:
            energy_axis = energy_axis_true.copy(name="energy")
        return cls([energy_axis_true, energy_axis], get_overlap_fraction(energy_axis, energy_axis_true))

    @classmethod
    def from_gauss(cls, energy_axis_true, energy_axis, sigma, bias, pdf_threshold=1e-6):
        """Gaussian in migra = E_reco / E_true around 1 + bias, integrated over each reco bin at the
        true-bin centre:  1/2 [erf((m_hi - 1 - bias)/(sqrt2 sigma)) - erf((m_lo - 1 - bias)/(sqrt2 sigma))]."""
        e_true = energy_axis_true.center[:, np.newaxis]
        lo = energy_axis.edges[:-1][np.newaxis, :] / e_true
        hi = energy_axis.edges[1:][np.newaxis, :] / e_true
        sigma = np.broadcast_to(np.asarray(sigma, dtype=float), (energy_axis_true.nbin,))[:, np.newaxis]
        bias = np.broadcast_to(np.asarray(bias, dtype=float), (energy_axis_true.nbin,))[:, np.newaxis]
        s = np.sqrt(2) * sigma
        pdf = (scipy.special.erf((hi - 1 - bias) / s) - scipy.special.erf((lo - 1 - bias) / s)) / 2
        pdf[pdf < pdf_threshold] = 0
        return cls([energy_axis_true, energy_axis], pdf)


class EDispKernelMap:
    """EDispKernelMap (irf/edisp/map.py): one kernel for the whole field, or -- ``data`` (nT, nR, ny, nx) on the image
    geometry ``geom`` -- one per IRF pixel; ``get_edisp_kernel(position)`` then returns the nearest pixel's."""

    def __init__(self, kernel=None, data=None, geom=None, axes=None):
        self.kernel = kernel
        self.data = None if data is None else np.asarray(data, dtype=float)
        self.geom = None if geom is None else geom.to_image()
        self.axes = axes  # [energy_axis_true, energy_axis] of a map with a spatial grid
        if self.data is not None:
            if self.geom is None or axes is None or self.data.ndim != 4:
                raise ValueError("an edisp kernel map with spatial axes needs data (nT, nR, ny, nx), geom and axes")
            if self.data.shape != (axes[0].nbin, axes[1].nbin, self.geom.npix[1], self.geom.npix[0]):
                raise ValueError("edisp kernel map data do not match the axes / image geometry")

    @property
    def has_single_spatial_bin(self):
        return self.data is None

    @classmethod
    def from_diagonal_response(cls, energy_axis, energy_axis_true, geom=None):
        return cls(EDispKernel.from_diagonal_response(energy_axis_true, energy_axis))

    @classmethod
    def from_gauss(cls, energy_axis, energy_axis_true, sigma, bias, pdf_threshold=1e-6, geom=None):
        return cls(EDispKernel.from_gauss(energy_axis_true, energy_axis, sigma, bias, pdf_threshold))

    @classmethod
    def from_edisp_kernel(cls, edisp, geom=None):
        return cls(edisp)

    def get_edisp_kernel(self, position=None, energy_axis=None):
        """irf/edisp/map.py:369-401 (``to_region_nd_map`` of a point: the nearest IRF pixel)."""
        if self.data is not None:
            if energy_axis is not None and not energy_axis == self.axes[1]:
                raise ValueError("EDispKernelMap: reco energy axis does not match the dataset geometry")
            if position is None:
                c = self.geom.center_skydir
                position = (c.lon, c.lat)
            lon, lat = (position.lon, position.lat) if hasattr(position, "lon") else position
            px, py = _irf_pixel(self.geom, float(lon), float(lat))
            i = int(np.clip(np.floor(px + 0.5), 0, self.geom.npix[0] - 1))
            j = int(np.clip(np.floor(py + 0.5), 0, self.geom.npix[1] - 1))
            return EDispKernel(list(self.axes), self.data[:, :, j, i])
        if energy_axis is not None and not energy_axis == self.kernel.axes["energy"]:
            raise ValueError("EDispKernelMap: reco energy axis does not match the dataset geometry")
        return self.kernel
