"""Oracle: ``MapEvaluator`` / ``MapDataset.npred`` / ``Datasets.stat_sum`` restated on plain numpy arrays.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference/gammapy):

* ``MapEvaluator.update / needs_update / cutout_width / irf_position_changed``  datasets/evaluator.py:128-243, 465-481
* ``compute_flux_psf_convolved / compute_flux_spatial / compute_flux_spectral``  datasets/evaluator.py:267-337
* ``apply_exposure / apply_psf / apply_edisp``      datasets/evaluator.py:346-377, datasets/utils.py:66-84
* ``WcsNDMap.convolve`` (float32 image, float32 kernel, scipy FFT, mode "same")    maps/wcs/ndmap.py:920-1010
* ``SkyModel.contributes``                          modeling/models/cube.py:246-286
* ``MapDataset.npred / npred_signal / npred_background``   datasets/map.py:775-880
* ``FoVBackgroundModel.evaluate_geom``              modeling/models/cube.py:737-756
* ``WcsNDMap.stack``                                maps/wcs/ndmap.py:1138-1181
* ``MapDataset.fake``                               datasets/map.py:1538-1554
* ``Datasets.stat_sum``                             datasets/core.py:264-277

Conventions: exposure in m2 s, flux in cm-2 s-1 => npred = flux * exposure * 1e4 (evaluator.py:351).
The static IRF arrays (PSF kernel cube ``(nT, K, K)``, edisp matrix ``(nT, nR)``) are inputs.

``conv`` selects the convolution arithmetic: ``"fft32"`` is the reference's own
(float32 FFT); ``"exact64"`` (float64 FFT) and ``"direct64"`` (float64 direct sum) are the exact
variants used to measure the reference's own float32 noise floor.
"""
import numpy as np
import scipy.signal
from . import cash as _cash
from . import irf as _irf
from . import spatial as _spatial
from . import spectral as _spectral
from .wcs import DEG, NoOverlapError, angular_separation

PSF_CONTAINMENT = 0.999  # evaluator.py:16
CUTOUT_MARGIN = 0.1  # deg, evaluator.py:17


class Dataset:
    """Plain-array stand-in for the MapDataset attributes the path reads."""

    def __init__(self, geom, energy_edges, energy_true_edges, exposure, background=None, counts=None,
                 mask_safe=None, mask_fit=None, psf_kernel=None, edisp=None, energy_center=None, name="ds",
                 psf_map=None, rad_edges=None, edisp_map=None, irf_geom=None):
        self.geom = geom
        self.energy_edges = np.asarray(energy_edges, float)
        self.energy_true_edges = np.asarray(energy_true_edges, float)
        if energy_center is None:  # MapAxis log-interp centre: exp(0.5 (log lo + log hi))
            le = np.log(self.energy_edges)
            energy_center = np.exp(0.5 * (le[:-1] + le[1:]))
        self.energy_center = np.asarray(energy_center, float)
        self.exposure = exposure
        self.background = background
        self.counts = counts
        self.mask_safe = mask_safe
        self.mask_fit = mask_fit
        self.psf_kernel = psf_kernel
        self.edisp = edisp
        self.name = name
        # position-dependent IRFs (oracle/irf.py): PSF map (nT, nrad, ny, nx) on rad_edges, edisp map (nT, nR, ny, nx),
        # both on the IRF image grid irf_geom (oracle.irf.IrfGeom)
        self.psf_map, self.rad_edges, self.edisp_map, self.irf_geom = psf_map, rad_edges, edisp_map, irf_geom

    @property
    def mask(self):
        """datasets/core.py:73-81."""
        if self.mask_safe is not None and self.mask_fit is not None:
            return self.mask_safe * self.mask_fit
        if self.mask_fit is not None:
            return self.mask_fit
        return self.mask_safe

    @property
    def mask_image(self):
        """datasets/map.py:1040-1048: all-true image when there is no mask, else OR over energy of the mask."""
        mask = self.mask
        if mask is None:
            return np.ones(self.geom.data_shape, dtype=bool)
        return np.logical_or.reduce(np.asarray(mask).astype(bool), axis=0)

    @property
    def data_shape(self):
        return (len(self.energy_edges) - 1,) + self.geom.data_shape


def convolve_fft32(image, kernel):
    """maps/wcs/ndmap.py:961,984,1002-1010: float32 image, float32 kernel planes, FFT, 'same', float32 out."""
    image = image.astype(np.float32)
    kernel = kernel.astype(np.float32)
    out = np.empty((kernel.shape[0],) + image.shape, dtype=np.float32)
    for idx in range(kernel.shape[0]):
        out[idx] = scipy.signal.convolve(image, kernel[idx], method="fft", mode="same")
    return out


def convolve_exact64(image, kernel, method="fft"):
    """Exact variant: same float32-rounded operands, float64 arithmetic (noise-floor reference).

    ``method="fft"`` (float64 FFT, error ~1e-16 x peak) is the fast default; ``"direct"`` is the plain sum."""
    image = image.astype(np.float32).astype(np.float64)
    kernel = kernel.astype(np.float32).astype(np.float64)
    out = np.empty((kernel.shape[0],) + image.shape, dtype=np.float64)
    for idx in range(kernel.shape[0]):
        out[idx] = scipy.signal.convolve(image, kernel[idx], method=method, mode="same")
    return out


def model_contributes(spatial, mask_image, geom, margin=0.0):
    """modeling/models/cube.py:246-286 (mask already reduced to an image)."""
    try:
        width = 2 * _spatial.evaluation_radius(spatial) + CUTOUT_MARGIN + margin
        _, (ys, xs) = geom.cutout_info(_spatial.map_position(spatial), width)
        return bool(np.any(mask_image[ys, xs]))
    except (NoOverlapError, ValueError):
        return False


class Evaluator:
    """One per sky model, stateful like ``MapEvaluator`` (local evaluation mode, npred cache)."""

    def __init__(self, model, conv="fft32", use_cache=True):
        self.model = model
        self.conv = conv
        self.use_cache = use_cache
        self.contributes = True
        self.exposure = None
        self.geom = None
        self.slices = None
        self._cached_position = (0.0, 0.0)  # evaluator.py:93 (radians)
        self._cached_spatial = None
        self._flux_spatial = None
        self.cutout_width = None

    # -- state -----------------------------------------------------------------------------
    def _spatial_values(self):
        s = self.model["spatial"]
        return tuple(s[k] for k in sorted(s) if k != "type")

    def parameters_spatial_changed(self, reset=True):
        values = self._spatial_values()
        changed = self._cached_spatial != values
        if changed and reset:
            self._cached_spatial = values
        return changed

    def irf_position_changed(self):
        """evaluator.py:465-481."""
        s = self.model["spatial"]
        lon_cached, lat_cached = self._cached_position
        lon, lat = s["lon_0"] * DEG, s["lat_0"] * DEG
        separation = angular_separation(lon, lat, lon_cached, lat_cached)
        changed = separation > (_spatial.evaluation_radius(s) + CUTOUT_MARGIN) * DEG
        if changed:
            self._cached_position = (lon, lat)
        return bool(changed)

    def needs_update(self):
        """evaluator.py:128-144."""
        if not self.contributes:
            return False
        if self.exposure is None:
            return True
        if not self.parameters_spatial_changed(reset=False):
            return False
        return self.irf_position_changed()

    def update(self, ds):
        """evaluator.py:174-243."""
        s = self.model["spatial"]
        self.psf = ds.psf_kernel
        self.edisp = ds.edisp
        position = _spatial.map_position(s)  # the dataset's maps are in the map frame
        if ds.edisp_map is not None:  # evaluator.py:196-204: edisp.get_edisp_kernel(position=self.position)
            self.edisp = _irf.edisp_kernel_at(ds.edisp_map, ds.irf_geom, position)
        if ds.psf_map is not None:    # evaluator.py:206-227: psf.get_psf_kernel(position, geom, containment=0.999)
            self.psf = _irf.psf_kernel_at(ds.psf_map, ds.rad_edges, ds.irf_geom, position, ds.geom.binsz,
                                          tuple(ds.geom.center_skydir), proj=ds.geom.proj)
            self.n_irf_updates = getattr(self, "n_irf_updates", 0) + 1
        psf_width = 0.0
        if self.psf is not None:
            psf_width = ds.geom.binsz * self.psf.shape[-1]  # np.max(kernel geom width) = cdelt * npix
        self.cutout_width = psf_width + 2 * (_spatial.evaluation_radius(s) + CUTOUT_MARGIN)
        self.contributes = model_contributes(s, ds.mask_image, ds.geom, margin=psf_width)
        if self.contributes:
            self.geom, self.slices = ds.geom.cutout_info(position, self.cutout_width, mode="trim", odd_npix=True)
            ys, xs = self.slices
            self.exposure = ds.exposure[:, ys, xs]
        self._flux_spatial = None
        self._cached_spatial = None

    # -- computation ---------------------------------------------------------------------------
    def compute_flux_spatial(self):
        """evaluator.py:282-325."""
        if self.parameters_spatial_changed() or not self.use_cache or self._flux_spatial is None:
            value = _spatial.integrate_geom(self.model["spatial"], self.geom)
            if self.psf is not None and self.model.get("apply_irf", {}).get("psf", True):
                if self.conv == "fft32":
                    value = convolve_fft32(value, self.psf)
                elif self.conv == "direct64":
                    value = convolve_exact64(value, self.psf, method="direct")
                else:
                    value = convolve_exact64(value, self.psf)
            else:
                value = value[np.newaxis]
            self._flux_spatial = value
        return self._flux_spatial

    def compute_flux_spectral(self, ds):
        """evaluator.py:327-337."""
        e = ds.energy_true_edges
        return _spectral.integral(self.model["spectral"], e[:-1], e[1:]).reshape((-1, 1, 1))

    def compute_npred(self, ds):
        """evaluator.py:379-417, standard sequence flux_psf_convolved -> exposure -> edisp."""
        spec = self.model["spectral"]
        if spec.get("amplitude", None) == 0:
            return np.zeros((len(ds.energy_edges) - 1,) + self.geom.data_shape)
        value = self.compute_flux_spectral(ds)
        value = value * self.compute_flux_spatial()  # float64 * float32 -> float64
        npred = (value * self.exposure) * 1e4  # (flux.quantity * exposure.quantity).to_value("")
        # apply_edisp (datasets/utils.py:66-84)
        data = np.moveaxis(npred, 0, -1)
        shape_space = data.shape[:-1]
        data_2d = data.reshape(-1, data.shape[-1])
        out = (data_2d @ self.edisp).reshape(shape_space + (self.edisp.shape[-1],))
        return np.moveaxis(out, -1, 0)


def background_factor(bkg_model, energy_center):
    """FoVBackgroundModel.evaluate_geom (cube.py:737-756) with PowerLawNormSpectralModel (spectral.py:1159-1162)."""
    spec = bkg_model["spectral"]
    return _spectral.plnorm_evaluate(energy_center, spec.get("tilt", 0.0), spec["norm"],
                                     spec.get("reference", 1.0)).reshape((-1, 1, 1))


class DatasetEvaluator:
    """``MapDataset`` model side: evaluators + npred + stat, with the reference's caching state."""

    def __init__(self, ds, models, conv="fft32"):
        self.ds = ds
        self.conv = conv
        self.set_models(models)

    def set_models(self, models):
        self.models = models
        self.bkg_model = None
        self.evaluators = {}
        for m in models:
            if m.get("type") == "fov-bkg":
                self.bkg_model = m
            else:
                self.evaluators[m["name"]] = Evaluator(m, conv=self.conv)

    def npred_signal(self, per_model=False):
        """datasets/map.py:824-880."""
        ds = self.ds
        total = np.zeros(ds.data_shape, dtype=float)
        parts = {}
        for name, ev in self.evaluators.items():
            if ev.needs_update():
                ev.update(ds)
            if ev.contributes:
                npred = ev.compute_npred(ds)
                data = npred.astype(total.dtype)
                not_finite = ~np.isfinite(data)
                if np.any(not_finite):
                    data = data.copy()
                    data[not_finite] = 0
                ys, xs = ev.slices
                total[:, ys, xs] += data
                if per_model:
                    part = np.zeros(ds.data_shape, dtype=float)
                    part[:, ys, xs] += data
                    parts[name] = part
        return (total, parts) if per_model else total

    def npred_background(self):
        """datasets/map.py:790-813."""
        ds = self.ds
        if self.bkg_model is not None and ds.background is not None:
            return ds.background * background_factor(self.bkg_model, ds.energy_center)
        return ds.background

    def npred(self):
        """datasets/map.py:775-788."""
        total = self.npred_signal()
        if self.ds.background is not None:
            total = total + self.npred_background()
        total[total < 0.0] = 0
        return total

    def stat_sum(self, use_reference=True):
        """datasets/core.py:88-90 -> stats/fit_statistics.py:281-293."""
        return _cash.stat_sum_dataset(self.ds.counts, self.npred(), self.ds.mask, use_reference=use_reference)

    def fake(self, random_state=0):
        """datasets/map.py:1538-1554 (legacy RandomState, utils/random/utils.py:81-92)."""
        rs = np.random.RandomState(random_state) if not isinstance(random_state, np.random.RandomState) else random_state
        npred = self.npred()
        data = np.nan_to_num(npred, copy=True, nan=0.0, posinf=0.0, neginf=0.0)
        self.ds.counts = rs.poisson(data).astype("float")
        return self.ds.counts


def stat_sum(dataset_evaluators, use_reference=True):
    """Datasets.stat_sum (datasets/core.py:264-277), no priors."""
    total = 0.0
    for de in dataset_evaluators:
        total += de.stat_sum(use_reference=use_reference)
    return total
