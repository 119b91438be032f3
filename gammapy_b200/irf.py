"""Static IRF producers: the PSF kernel cube ``(nT, K, K)`` and the edisp matrix ``(nT, nR)``.

These are *inputs* of the likelihood hot path, built once on the host with numpy.
Reference: irf/psf/map.py (PSFMap.from_gauss :397-448, get_psf_kernel :249-337, _psf_upsampling_factor
:23-37), irf/psf/kernel.py (PSFKernel :12-126), irf/psf/core.py (containment_radius :39-87),
irf/edisp/map.py (EDispKernelMap, get_overlap_fraction :11-20), irf/edisp/kernel.py (EDispKernel).

PSFMap here is position independent (one radial profile per true energy, as ``from_gauss`` makes it).
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


class PSFMap:
    """Position-independent PSF: radial profile ``psf(energy_true, rad)`` in sr-1 (irf/psf/map.py)."""

    energy_name = "energy_true"

    def __init__(self, energy_axis_true, rad_axis, data):
        self.energy_axis_true = energy_axis_true
        self.rad_axis = rad_axis
        self.data = np.asarray(data, dtype=float)  # (nT, nrad) at rad_axis.center
        self.has_single_spatial_bin = True

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
        """irf/psf/map.py:249-337 for a position-independent PSF."""
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
    """Position-independent EDispKernelMap (irf/edisp/map.py): one kernel for the whole field."""

    def __init__(self, kernel):
        self.kernel = kernel

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
        """irf/edisp/map.py:369-401."""
        if energy_axis is not None and not energy_axis == self.kernel.axes["energy"]:
            raise ValueError("EDispKernelMap: reco energy axis does not match the dataset geometry")
        return self.kernel
