"""``MapDataset`` / ``Datasets`` with the reference's API, evaluated on the GPU.

Reference: datasets/map.py (MapDataset :410, create :947-1024, npred :775-880, fake :1538-1554),
datasets/core.py (Dataset.mask :73-81, stat_sum_likelihood :88-90, Datasets.stat_sum :264-277).
Maps are numpy on the host (user-facing, mutable); the first likelihood evaluation mirrors them into HBM.
After an in-place edit of map data call ``reset_data_cache()`` (re-assigning an attribute is detected).
"""
import itertools

import numpy as np

from .irf import RAD_AXIS_DEFAULT, EDispKernelMap, PSFKernel, PSFMap
from .maps import Map, MapAxis
from .modeling.models import FoVBackgroundModel, Models, SkyModel
from .modeling.parameter import Parameters
from .utils import get_random_state

__all__ = ["MapDataset", "MapDatasetOnOff", "Datasets"]

_ds_counter = itertools.count()
PSF_CONTAINMENT = 0.999  # datasets/evaluator.py:16
PSF_MAX_RADIUS = None    # datasets/evaluator.py:15


class MapDataset:
    """Binned 3D dataset with Cash statistic (datasets/map.py:410)."""

    tag = "MapDataset"
    stat_type = "cash"

    def __init__(self, models=None, counts=None, exposure=None, background=None, psf=None, edisp=None,
                 mask_safe=None, mask_fit=None, gti=None, meta_table=None, name=None, meta=None, stat_type="cash",
                 storage="float32"):
        if stat_type not in ("cash", "cash_weighted"):
            raise NotImplementedError("only stat_type 'cash' and 'cash_weighted' are on the accelerated path")
        self.stat_type = stat_type
        self.storage = storage
        self._name = name if name is not None else f"dataset-{next(_ds_counter):04d}"
        self._device = None
        self._engine = None
        self._models_version = 0
        self._source_names = []
        self._psf_kernel_cache = None
        self._counts = counts
        self._exposure = exposure
        self._background = background
        self._psf = psf
        self._edisp = edisp
        self._mask_safe = mask_safe
        self._mask_fit = mask_fit
        self.gti = gti
        self.meta_table = meta_table
        self.meta = meta
        self._models = None
        self.models = models

    @property
    def storage(self):
        """HBM layout of the counts and the mask: ``"float32"`` (counts float32, mask one byte per bin) or
        ``"compact"`` (counts uint16 -- refused above 65535 --, mask one bit per bin): the format of stacked fits with
        hundreds of datasets per GPU (BASELINE.json configs[4]).  Exposure maps that are the same ``Map`` object in
        several datasets are uploaded once either way."""
        return self._storage

    @storage.setter
    def storage(self, value):
        if value not in ("float32", "compact"):
            raise ValueError("storage must be 'float32' or 'compact'")
        if value == "compact" and self.stat_type == "cash_weighted":
            raise NotImplementedError("float mask weights cannot be bit-packed: 'compact' needs stat_type='cash'")
        if getattr(self, "_storage", value) != value and getattr(self, "_device", None) is not None:
            self._invalidate()
        self._storage = value

    # -- map attributes: setters invalidate the device mirror ------------------------------------------------
    def _invalidate(self, counts_only=False):
        if self._device is not None and counts_only and self._counts is not None:
            self._device.refresh_counts()
            return
        if self._device is not None:
            self._device.close()
        self._device = None
        self._engine = None
        self._psf_kernel_cache = None

    name = property(lambda self: self._name)
    counts = property(lambda self: self._counts)
    exposure = property(lambda self: self._exposure)
    background = property(lambda self: self._background)
    psf = property(lambda self: self._psf)
    edisp = property(lambda self: self._edisp)
    mask_safe = property(lambda self: self._mask_safe)
    mask_fit = property(lambda self: self._mask_fit)

    @counts.setter
    def counts(self, value):
        had = self._counts is not None
        self._counts = value
        self._invalidate(counts_only=had and value is not None)

    @exposure.setter
    def exposure(self, value):
        self._exposure = value
        self._invalidate()

    @background.setter
    def background(self, value):
        self._background = value
        self._invalidate()

    @psf.setter
    def psf(self, value):
        self._psf = value
        self._invalidate()

    @edisp.setter
    def edisp(self, value):
        self._edisp = value
        self._invalidate()

    @mask_safe.setter
    def mask_safe(self, value):
        self._mask_safe = value
        self._invalidate()

    @mask_fit.setter
    def mask_fit(self, value):
        self._mask_fit = value
        self._invalidate()

    def reset_data_cache(self):
        """Drop the HBM mirror (and evaluator caches): call after editing map data in place."""
        self._invalidate()

    def _data_fingerprint(self):
        """Identity of the map arrays the HBM mirror was built from.  ``DeviceDataset`` keeps references to the
        arrays it uploaded (``_held``), so an id cannot be reused by another array while the mirror lives."""
        maps = (self._counts, self._exposure, self._background, self._mask_safe, self._mask_fit)
        return tuple(id(m.data) if m is not None else None for m in maps)

    def _data_arrays(self):
        maps = (self._counts, self._exposure, self._background, self._mask_safe, self._mask_fit)
        return tuple(m.data if m is not None else None for m in maps)

    @property
    def _geom(self):
        """Main analysis geometry (datasets/map.py:701-716)."""
        for m in (self._counts, self._background, self._mask_safe, self._mask_fit):
            if m is not None:
                return m.geom
        raise ValueError("Either 'counts', 'background', 'mask_fit' or 'mask_safe' must be defined.")

    @property
    def geoms(self):
        out = {"geom": self._geom}
        if self._exposure is not None:
            out["geom_exposure"] = self._exposure.geom
        return out

    @property
    def data_shape(self):
        return self._geom.data_shape

    @property
    def mask(self):
        """Combined fit and safe mask (datasets/core.py:73-81)."""
        if self.mask_safe is not None and self.mask_fit is not None:
            return self.mask_safe * self.mask_fit
        if self.mask_fit is not None:
            return self.mask_fit
        return self.mask_safe

    @property
    def mask_image(self):
        """datasets/map.py:1040-1048."""
        mask = self.mask
        if mask is None:
            m = Map.from_geom(self._geom.to_image(), dtype=bool)
            m.data |= True
            return m
        return mask.reduce_over_axes(func=np.logical_or)

    # -- models ------------------------------------------------------------------------------------------------
    @property
    def models(self):
        return self._models

    @models.setter
    def models(self, models):
        """datasets/map.py:674-694: new evaluators, i.e. every cache is dropped."""
        if models is not None:
            models = Models(models).select(datasets_names=self.name)
            for m in models:
                if not isinstance(m, (SkyModel, FoVBackgroundModel)):
                    raise NotImplementedError(f"{type(m).__name__} is outside the accelerated path")
        self._models = models
        self._models_version += 1
        self._engine = None

    @property
    def background_model(self):
        if self._models is None:
            return None
        for m in self._models:
            if isinstance(m, FoVBackgroundModel):
                return m
        return None

    @property
    def parameters(self):
        return self._models.parameters if self._models is not None else Parameters([])

    # -- IRF reduction (static inputs of the path) ---------------------------------------------------------------
    def _psf_kernel_for_evaluators(self):
        """``PSFMap.get_psf_kernel(geom, containment=0.999)`` (evaluator.py:221-226) for a position-independent PSF."""
        if self._psf is None:
            return None
        if isinstance(self._psf, PSFKernel):
            return self._psf
        if self._psf_kernel_cache is None:
            self._psf_kernel_cache = self._psf.get_psf_kernel(
                geom=self._exposure.geom, containment=PSF_CONTAINMENT, max_radius=PSF_MAX_RADIUS)
        return self._psf_kernel_cache

    # -- device side ----------------------------------------------------------------------------------------------
    def _device_dataset(self, ctx):
        from .engine import DeviceDataset

        if self._exposure is None:
            raise ValueError("MapDataset needs an exposure map to evaluate models")
        if self._device is not None and self._device._uploaded != self._data_fingerprint():
            self._invalidate()
        if self._device is None:
            self._device = DeviceDataset(self, ctx)
        return self._device

    def _get_engine(self):
        from .engine import LikelihoodEngine

        stale = self._engine is None or self._engine._signature != LikelihoodEngine.signature([self])
        if not stale and self._device._uploaded != self._data_fingerprint():
            stale = True
        if stale:
            self._engine = LikelihoodEngine([self])
        return self._engine

    # -- predictions ------------------------------------------------------------------------------------------------
    def _npred_map(self, data):
        return Map.from_geom(self._geom, data=data, dtype=float)

    def npred(self):
        """Total predicted counts (datasets/map.py:775-788)."""
        return self._npred_map(self._get_engine().npred(0, include_background=True))

    def npred_signal(self, model_names=None, stack=True):
        """Predicted signal counts (datasets/map.py:824-880)."""
        eng = self._get_engine()
        names = self._source_names
        if model_names is not None:
            if isinstance(model_names, str):
                model_names = [model_names]
            for n in model_names:
                if n not in names:
                    raise KeyError(n)
        if stack:
            enable = None if model_names is None else [n in model_names for n in names]
            return self._npred_map(eng.npred(0, source_enable=enable, include_background=False))
        selected = names if model_names is None else model_names
        cubes = []
        for n in selected:
            enable = [m == n for m in names]
            cubes.append(eng.npred(0, source_enable=enable, include_background=False))
        label_axis = MapAxis(np.arange(len(selected) + 1) - 0.5, name="models")
        label_axis.labels = list(selected)
        geom = self._geom.to_cube([label_axis])
        return Map.from_geom(geom, data=np.array(cubes), dtype=float)

    def npred_background(self):
        """Predicted background counts (datasets/map.py:790-813) = npred with every source switched off."""
        if self._background is None:
            return None
        eng = self._get_engine()
        data = eng.npred(0, source_enable=[False] * len(self._source_names), include_background=True)
        return self._npred_map(data)

    def stat_sum_likelihood(self):
        """Cash sum given the current parameters (datasets/core.py:88-90, fit_statistics.py:281-293)."""
        if self._counts is None:
            raise ValueError("MapDataset has no counts")
        return self._get_engine().stat_sum()

    stat_sum = stat_sum_likelihood

    # -- stacking (data reduction before a fit; host arithmetic) ----------------------------------------------------
    @classmethod
    def from_geoms(cls, geom, geom_exposure=None, geom_psf=None, geom_edisp=None, name=None, **kwargs):
        """Zero-filled dataset on the given geometries (datasets/map.py:883-945): counts, background, an all-False
        ``mask_safe`` and, when its geometry is given, the exposure; IRFs are taken over by the first ``stack``."""
        del geom_psf, geom_edisp
        kwargs.setdefault("counts", Map.from_geom(geom, unit=""))
        kwargs.setdefault("background", Map.from_geom(geom, unit=""))
        kwargs.setdefault("mask_safe", Map.from_geom(geom, unit="", dtype=bool))
        if geom_exposure is not None:
            kwargs.setdefault("exposure", Map.from_geom(geom_exposure, unit="m2 s"))
        return cls(name=name, **kwargs)

    @property
    def mask_safe_image(self):
        """``mask_safe`` reduced over energy with a logical or (datasets/map.py:1050-1056), or None."""
        if self.mask_safe is None:
            return None
        return np.asarray(self.mask_safe.data, bool).any(axis=0)

    def _background_for_stacking(self):
        """``npred_background()`` (datasets/map.py:790-813): the background map itself without a background model,
        else the GPU's evaluation of it."""
        if self.background_model is None:
            return np.asarray(self._background.data, float)
        return np.asarray(self.npred_background().data, float)

    def _irf_weight(self):
        """Exposure weight per true-energy bin for the IRF average: the exposure summed over the safe image pixels
        (the reference carries an exposure map inside every IRF map, filled by the makers with the exposure at the IRF
        pixel; for the position-independent IRFs handled here that is one number per true-energy bin)."""
        data = np.asarray(self._exposure.data, float)
        safe = self.mask_safe_image
        if safe is not None:
            data = data * safe[np.newaxis]
        return data.sum(axis=(1, 2))

    def stack(self, other, nan_to_num=True):
        """Stack ``other`` onto this dataset in place (datasets/map.py:1104-1210): its counts, exposure and predicted
        background enter where its ``mask_safe`` is true, the safe and fit masks are or-ed, position-independent PSF /
        edisp are averaged with exposure weights.  The background model of ``self`` is folded into the background map
        (as the reference does through ``npred_background()``), so assign the models of the stacked dataset afterwards."""
        if isinstance(other, MapDatasetOnOff) != isinstance(self, MapDatasetOnOff):
            raise TypeError("Incompatible types for stacking: a MapDataset and a MapDatasetOnOff")
        o_safe = None if other.mask_safe is None else np.asarray(other.mask_safe.data, bool)
        if o_safe is not None and not o_safe.any():
            return

        def add(target, values, weights):
            values = np.asarray(values, float)
            if nan_to_num:
                values = np.where(np.isfinite(values), values, 0.0)
            return target + (values if weights is None else values * weights)

        w_self = self._irf_weight() if self._exposure is not None else None
        w_other = other._irf_weight() if other._exposure is not None else None
        if self._counts is not None and other._counts is not None:
            self.counts = Map.from_geom(self._geom, data=add(np.asarray(self._counts.data, float), other._counts.data, o_safe),
                                        dtype=float)
        if self._exposure is not None and other._exposure is not None:
            o_img = None if o_safe is None else o_safe.any(axis=0)[np.newaxis]
            self.exposure = Map.from_geom(self._exposure.geom, unit="m2 s",
                                          data=add(np.asarray(self._exposure.data, float), other._exposure.data, o_img))
        if self._background is not None and other._background is not None:
            stacked = add(self._background_for_stacking(), other._background_for_stacking(), o_safe)
            self._models = None if self._models is None else Models(
                [m for m in self._models if not isinstance(m, FoVBackgroundModel)]) or None
            self.background = Map.from_geom(self._geom, data=stacked, dtype=float)
        for attr in ("psf", "edisp"):
            mine, theirs = getattr(self, attr), getattr(other, attr)
            if theirs is None:
                continue
            if mine is None or w_self is None or not np.any(w_self):
                setattr(self, attr, theirs)   # the first dataset stacked onto an empty one brings its IRFs
                continue
            if not (mine.has_single_spatial_bin and theirs.has_single_spatial_bin):
                raise NotImplementedError(f"stacking {attr} maps with a spatial grid is outside the accelerated path")
            setattr(self, attr, _average_irf(attr, mine, theirs, w_self, w_other))
        if self.mask_safe is not None and o_safe is not None:
            self.mask_safe = Map.from_geom(self._geom, data=np.asarray(self.mask_safe.data, bool) | o_safe, dtype=bool)
        if other.mask_fit is not None:
            o_fit = np.asarray(other.mask_fit.data, bool)
            mine = o_fit if self.mask_fit is None else (np.asarray(self.mask_fit.data, bool) | o_fit)
            self.mask_fit = Map.from_geom(self._geom, data=mine.copy(), dtype=bool)
        self.models = self._models  # (new evaluators: every cache is dropped)

    def copy(self, name=None):
        """Deep copy of the data under a new name, without models (datasets/core.py:96-114); the HBM mirror and the
        engines are not copied -- the copy registers itself with the library when it is first evaluated."""
        import copy as _copy

        kwargs = {"name": name, "storage": self.storage, "meta": _copy.deepcopy(self.meta),
                  "psf": _copy.deepcopy(self.psf), "edisp": _copy.deepcopy(self.edisp)}
        if not isinstance(self, MapDatasetOnOff):
            kwargs["stat_type"] = self.stat_type
        for attr in ("counts", "exposure", "background", "mask_safe", "mask_fit", "counts_off", "acceptance", "acceptance_off"):
            m = getattr(self, attr, None)
            if m is not None:
                kwargs[attr] = m.copy()
        return type(self)(**kwargs)

    def cutout(self, position, width, mode="trim", name=None):
        """The dataset on a sub-region (datasets/map.py:2040-2094): every map cut with the integer rules of
        ``WcsGeom.cutout`` (astropy ``Cutout2D``), the models kept.  Position-independent IRFs are shared; IRF maps on a
        spatial grid stay whole -- they carry their own geometry, the evaluators look them up by sky position."""
        kwargs = {"name": name, "stat_type": self.stat_type, "models": self.models, "psf": self.psf, "edisp": self.edisp}
        for attr in ("counts", "exposure", "background", "mask_safe", "mask_fit"):
            m = getattr(self, attr)
            if m is not None:
                cut = m.cutout(position, width, mode=mode)
                kwargs[attr] = Map.from_geom(cut.geom, data=cut.data, unit=getattr(m, "unit", ""), dtype=cut.data.dtype)
        return MapDataset(**kwargs)

    def to_masked(self, name=None, nan_to_num=True):
        """A copy with counts, background and exposure multiplied by the safe mask (datasets/map.py:1958-1986)."""
        geoms = dict(self.geoms)
        out = type(self).from_geoms(geoms["geom"], geom_exposure=geoms.get("geom_exposure"), name=name)
        out.stack(self, nan_to_num=nan_to_num)
        return out

    # -- what one looks at after a fit: all host arithmetic on the GPU's npred cubes ---------------------------------
    @property
    def excess(self):
        """Observed excess: counts - background (datasets/map.py:669-672)."""
        return Map.from_geom(self._geom, data=np.asarray(self._counts.data, float) - np.asarray(self._background.data, float),
                             dtype=float)

    def residuals(self, method="diff", **kwargs):
        """Residual map (datasets/map.py:1210-1255, core.py:117-129): ``diff`` = data - model, ``diff/model``,
        ``diff/sqrt(model)``; both sides multiplied by the mask, masked bins NaN.  Smoothing keywords are not supported."""
        if kwargs:
            raise NotImplementedError("residuals(): smoothing (Map.smooth keywords) is outside the accelerated path")
        npred = np.asarray(self.npred().data, float)
        counts = np.asarray(self._counts.data, float).copy()
        mask = None if self.mask is None else np.asarray(self.mask.data, bool)
        if mask is not None:
            npred, counts = npred * mask, counts * mask
        with np.errstate(invalid="ignore", divide="ignore"):
            if method == "diff":
                res = counts - npred
            elif method == "diff/model":
                res = (counts - npred) / npred
            elif method == "diff/sqrt(model)":
                res = (counts - npred) / np.sqrt(npred)
            else:
                raise AttributeError(f"Invalid method: {method!r} for computing residuals")
        if mask is not None:
            res[~mask] = np.nan
        return Map.from_geom(self._geom, data=res, dtype=float)

    def info_dict(self, in_safe_data_range=True):
        """Summary statistics summed over energy (datasets/map.py:1791-1900): counts, background, excess and its
        significance (``CashCountsStatistic`` of the summed numbers, stats/counts_statistic.py:250-288), the npred sums,
        the exposure range and the fit statistic.  Times (livetime, ontime, rates) are not carried by this package."""
        info = {"name": self.name}
        mask = np.asarray(self.mask_safe.data, bool) if (self.mask_safe is not None and in_safe_data_range) else slice(None)
        counts, background, excess, sqrt_ts = 0, np.nan, np.nan, np.nan
        npred_background = None if self._background is None else np.asarray(self.npred_background().data, float)
        if self._counts is not None:
            counts = float(np.asarray(self._counts.data, float)[mask].sum())
            if self._background is not None:
                background = float(np.asarray(self._background.data, float)[mask].sum())
                from .stats import cash_sqrt_ts

                mu_bkg = float(npred_background[mask].sum())
                excess = counts - mu_bkg
                sqrt_ts = float(cash_sqrt_ts(counts, mu_bkg))
        info.update(counts=int(counts), excess=float(excess), sqrt_ts=sqrt_ts, background=float(background))
        info["npred"] = float(self.npred().data[mask].sum()) if (self.models or not np.isnan(background)) else np.nan
        info["npred_background"] = float(npred_background[mask].sum()) if npred_background is not None else np.nan
        has_sources = self.models is not None and any(not isinstance(m, FoVBackgroundModel) for m in self.models)
        info["npred_signal"] = float(self.npred_signal().data[mask].sum()) if has_sources else np.nan
        exposure_min = exposure_max = np.nan
        if self.exposure is not None:
            data = np.asarray(self.exposure.data, float)
            sel = data > 0
            if self.mask_safe is not None:
                sel = sel & np.asarray(self.mask_safe.data, bool).any(axis=0)[np.newaxis]
            if not sel.any():
                sel = slice(None)
            exposure_min, exposure_max = float(data[sel].min()), float(data[sel].max())
        info.update(exposure_min=exposure_min, exposure_max=exposure_max)
        info["n_bins"] = int(np.asarray(self._counts.data).size) if self._counts is not None else 0
        info["n_fit_bins"] = int(np.sum(self.mask.data)) if self.mask is not None else 0
        info["stat_type"] = self.stat_type
        info["stat_sum"] = float(self.stat_sum()) if (self._counts is not None and self.models is not None) else np.nan
        return info

    def stat_array(self):
        """Per-bin Cash (stats/fit_statistics.py:29-79, 295-298, 323-329) from the GPU npred cube."""
        n_on = np.asarray(self._counts.data, dtype=float)
        mu_on = self.npred().data
        mu_on = np.where(mu_on <= 1e-25, 1e-25, mu_on)
        with np.errstate(divide="ignore", invalid="ignore"):
            stat = 2 * (mu_on - n_on * np.log(mu_on))
        if self.stat_type == "cash_weighted" and self.mask is not None:
            stat = stat * np.asarray(self.mask.data, dtype=float)
        return stat

    def fake(self, random_state="random-seed"):
        """Poisson counts from npred with the legacy RandomState (datasets/map.py:1538-1554)."""
        random_state = get_random_state(random_state)
        npred = self.npred()
        data = np.nan_to_num(npred.data, copy=True, nan=0.0, posinf=0.0, neginf=0.0)
        npred.data = random_state.poisson(data).astype("float")
        self.counts = npred

    def evaluator_state(self, model_name):
        """Cutout / cache state of one model's evaluator (MapEvaluator, datasets/evaluator.py:174-243)."""
        owner = getattr(self._device, "models_owner", None) if self._device is not None else None
        if owner is not None and owner._signature == owner.signature(owner.datasets):
            return owner.source_state(owner.datasets.index(self), self._source_names.index(model_name))
        eng = self._get_engine()
        return eng.source_state(0, self._source_names.index(model_name))

    # -- on-disk format (datasets/map.py:1556-1780; gammapy_b200/fitsio.py) ---------------------------------------------
    def to_hdulist(self):
        from . import fitsio

        return fitsio.dataset_to_hdus(self)

    @classmethod
    def from_hdulist(cls, hdulist, name=None):
        from . import fitsio

        return fitsio.dataset_from_hdus(hdulist, name=name)

    def write(self, filename, overwrite=False):
        """Write the maps and IRFs as one GADF FITS file (``MapDataset.write``, datasets/map.py:1741-1780); models are
        not part of the file, as in the reference."""
        from . import fitsio

        fitsio.write_hdus(str(filename), self.to_hdulist(), overwrite=overwrite)

    @classmethod
    def read(cls, filename, name=None):
        """``MapDataset.read`` (datasets/map.py:1702-1739)."""
        from . import fitsio

        return cls.from_hdulist(fitsio.read_hdus(str(filename)), name=name)

    def to_dict(self):
        """Entry of the datasets index file (datasets/core.py:67-71)."""
        return {"name": self.name, "type": self.tag, "filename": self.name.replace(" ", "_") + ".fits"}

    def evaluator_irf(self, model_name, with_kernel=True):
        """IRFs one model's evaluator looked up at its position (datasets with a PSFMap / EDispKernelMap on a spatial
        grid; MapEvaluator.update, datasets/evaluator.py:196-227): dict with the PSF kernel ``(nT, K, K)``, its side
        ``k``, the IRF pixel ``edisp_pixel`` (iy * nx + ix) of the edisp matrix and the number of look-ups."""
        self._get_engine()
        kernel, k, pix, n = self._device.source_irf(self._source_names.index(model_name), with_kernel)
        return {"psf_kernel": kernel, "k": k, "edisp_pixel": pix, "n_updates": n}

    # -- construction --------------------------------------------------------------------------------------------------
    @classmethod
    def create(cls, geom, energy_axis_true=None, migra_axis=None, rad_axis=None, binsz_irf=0.2,
               reference_time="2000-01-01", name=None, meta_table=None, reco_psf=False, **kwargs):
        """Zero-filled dataset (datasets/map.py:947-1024): float32 counts / background / exposure, boolean
        mask_safe (all False, as the reference), Gaussian-free default IRFs: PSF none, edisp diagonal."""
        if migra_axis is not None or reco_psf:
            raise NotImplementedError("EDispMap (migra) and RecoPSFMap are outside the accelerated path")
        energy_axis = geom.axes["energy"]
        if energy_axis_true is None:
            energy_axis_true = energy_axis.copy(name="energy_true")
        if energy_axis_true.name != "energy_true":
            raise ValueError("energy_axis_true must be named 'energy_true'")
        geom_image = geom.to_image()
        geom_true = geom_image.to_cube([energy_axis_true])
        kwargs.setdefault("counts", Map.from_geom(geom, unit=""))
        kwargs.setdefault("background", Map.from_geom(geom, unit=""))
        kwargs.setdefault("exposure", Map.from_geom(geom_true, unit="m2 s"))
        kwargs.setdefault("mask_safe", Map.from_geom(geom, unit="", dtype=bool))
        kwargs.setdefault("edisp", EDispKernelMap.from_diagonal_response(energy_axis, energy_axis_true))
        kwargs.setdefault("psf", None)
        del rad_axis, binsz_irf, reference_time
        return cls(name=name, meta_table=meta_table, **kwargs)

    def __repr__(self):
        return f"MapDataset(name={self.name!r}, geom={self._geom!r}, models={None if self._models is None else self._models.names})"


class MapDatasetOnOff(MapDataset):
    """ON / OFF map dataset with the WStat statistic (datasets/map.py:2554-2700, stats/fit_statistics.py:332-361):
    the background is not modelled but measured -- ``counts_off`` in an OFF region of relative acceptance
    ``acceptance_off / acceptance`` -- and profiled out bin by bin.  ``npred_signal`` comes from the fold kernel
    (no background term), the statistic from ``wstat_sum_kernel`` on the same device arrays."""

    tag = "MapDatasetOnOff"

    def __init__(self, counts_off=None, acceptance=None, acceptance_off=None, **kwargs):
        if kwargs.get("background") is not None:
            raise ValueError("MapDatasetOnOff takes no background model map: the OFF counts are the background")
        kwargs.setdefault("stat_type", "cash")  # for the parent's checks; this class overrides every stat method
        super().__init__(**kwargs)
        self.stat_type = "wstat"
        self.counts_off = counts_off
        geom = self._geom
        ones = np.ones(geom.data_shape)
        self.acceptance = acceptance if acceptance is not None else Map.from_geom(geom, data=ones.copy(), dtype=float)
        self.acceptance_off = (acceptance_off if acceptance_off is not None
                               else Map.from_geom(geom, data=ones.copy(), dtype=float))
        self._onoff_dev = None

    @classmethod
    def from_geoms(cls, geom, geom_exposure=None, geom_psf=None, geom_edisp=None, name=None, **kwargs):
        """Zero-filled ON/OFF dataset (datasets/map.py:2770-2830): also the OFF counts and both acceptances are zero."""
        for key in ("counts_off", "acceptance", "acceptance_off"):
            kwargs.setdefault(key, Map.from_geom(geom, unit="", dtype=float))
        kwargs.setdefault("counts", Map.from_geom(geom, unit=""))
        kwargs.setdefault("mask_safe", Map.from_geom(geom, unit="", dtype=bool))
        if geom_exposure is not None:
            kwargs.setdefault("exposure", Map.from_geom(geom_exposure, unit="m2 s"))
        return cls(name=name, **kwargs)

    def stack(self, other, nan_to_num=True):
        """Stack another ON/OFF dataset in place (datasets/map.py:2936-3006): ON and OFF counts and the ON acceptance
        add up under the other dataset's safe mask; the OFF acceptance is rescaled so that the stacked alpha is the
        OFF-count weighted mean, alpha = (alpha_1 OFF_1 + alpha_2 OFF_2) / (OFF_1 + OFF_2), with the run-averaged alpha
        where no OFF event was seen."""
        if not isinstance(other, MapDatasetOnOff):
            raise TypeError("Incompatible types for MapDatasetOnOff stacking")
        for d in (self, other):
            if d.counts_off is None or d.acceptance is None or d.acceptance_off is None:
                raise ValueError("Cannot stack incomplete MapDatasetOnOff.")

        def clean(a):
            a = np.asarray(a, float)
            return np.where(np.isfinite(a), a, 0.0) if nan_to_num else a

        w = 1.0 if other.mask_safe is None else np.asarray(other.mask_safe.data, bool)
        total_acceptance = clean(self.acceptance.data) + clean(other.acceptance.data) * w
        total_off = clean(self.counts_off.data) + clean(other.counts_off.data) * w
        total_alpha = clean(self.alpha.data * self.counts_off.data) + clean(other.alpha.data * other.counts_off.data) * w
        with np.errstate(divide="ignore", invalid="ignore"):
            acceptance_off = total_acceptance * total_off / total_alpha
            average_alpha = total_alpha.sum() / total_off.sum()
            is_zero = total_off == 0
            acceptance_off[is_zero] = total_acceptance[is_zero] / average_alpha
        geom = self._geom
        self.acceptance = Map.from_geom(geom, data=total_acceptance, dtype=float)
        self.acceptance_off = Map.from_geom(geom, data=acceptance_off, dtype=float)
        self.counts_off = Map.from_geom(geom, data=total_off, dtype=float)
        self._onoff_dev = None
        MapDataset.stack(self, other, nan_to_num=nan_to_num)

    @property
    def alpha(self):
        """Exposure ratio between the ON and the OFF region (datasets/map.py:2661-2677): acceptance / acceptance_off,
        0 where the OFF acceptance is 0 (np.nan_to_num of the division)."""
        with np.errstate(invalid="ignore", divide="ignore"):
            data = np.nan_to_num(np.asarray(self.acceptance.data, float) / np.asarray(self.acceptance_off.data, float))
        return Map.from_geom(self._geom, data=data, dtype=float)

    def npred_background(self):
        """Profile-likelihood background in the ON region, alpha * mu_bkg (datasets/map.py:2679-2700,
        stats/fit_statistics.py:221-236), from the GPU's npred_signal."""
        n_on, n_off = np.asarray(self._counts.data, float), np.asarray(self.counts_off.data, float)
        alpha, mu_sig = self.alpha.data, self.npred_signal().data
        c = alpha * (n_on + n_off) - (1 + alpha) * mu_sig
        d = np.sqrt(c ** 2 + 4 * alpha * (alpha + 1) * n_off * mu_sig)
        with np.errstate(invalid="ignore", divide="ignore"):
            mu_bkg = np.nan_to_num((c + d) / (2 * alpha * (alpha + 1)))
        return Map.from_geom(self._geom, data=alpha * mu_bkg, dtype=float)

    def _device_arrays(self):
        import torch

        eng = self._get_engine()
        key = (id(self._counts.data), id(self.counts_off.data), id(self.acceptance.data), id(self.acceptance_off.data),
               None if self.mask is None else id(self.mask.data))
        if self._onoff_dev is None or self._onoff_dev[0] != key:
            dev = eng.ctx.torch_device
            f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
            mask = None if self.mask is None else torch.from_numpy(
                np.ascontiguousarray(np.asarray(self.mask.data) != 0, dtype=np.uint8)).to(dev)
            self._onoff_dev = (key, f64(self._counts.data), f64(self.counts_off.data), f64(self.alpha.data), mask,
                               (self._counts.data, self.counts_off.data, self.acceptance.data, self.acceptance_off.data))
        return self._onoff_dev[1:5]

    def _wstat(self, return_array, values=None):
        from . import stats

        if self.counts_off is None:
            return (0.0, np.zeros(self._geom.data_shape)) if return_array else 0.0
        n_on, n_off, alpha, mask = self._device_arrays()
        mu_sig = self._get_engine().npred_device(0, values=values, include_background=False)
        return stats.wstat_sum_cuda(n_on, n_off, alpha, mu_sig, mask=mask, return_array=return_array)

    def _fit_engine(self):
        """What ``Fit`` needs (parameters, values, stat_sum of one or many parameter vectors) with WStat as the
        statistic: every vector is one npred_signal evaluation + one ``wstat_sum_kernel`` launch."""
        ds = self

        class _OnOffEngine:
            def __init__(self):
                self.inner = ds._get_engine()
                self.parameters = self.inner.parameters
                self.n_par = self.inner.n_par
                self.ctx = self.inner.ctx

            def values(self):
                return self.inner.values()

            def stat_sum(self, values=None):
                if values is None:
                    return ds._wstat(False)
                values = np.asarray(values, dtype=float)
                if values.ndim == 1:
                    return ds._wstat(False, values)
                return np.array([ds._wstat(False, v) for v in values])

        return _OnOffEngine()

    def stat_sum_likelihood(self):
        """``WStatFitStatistic.stat_sum_dataset`` (stats/fit_statistics.py:351-361)."""
        if self._counts is None:
            raise ValueError("MapDatasetOnOff has no counts")
        return self._wstat(False)

    stat_sum = stat_sum_likelihood

    def stat_array(self):
        """``WStatFitStatistic.stat_array_dataset`` (stats/fit_statistics.py:335-349)."""
        return self._wstat(True)[1]


def _parameters_with_prior(engine):
    """The engine's parameters that carry a prior; rescanned only after a prior was assigned somewhere."""
    from .modeling.parameter import Parameter

    cached = getattr(engine, "_prior_cache", None)
    if cached is None or cached[0] != Parameter.prior_epoch:
        cached = (Parameter.prior_epoch, [p for p in engine.parameters if p.prior is not None])
        engine._prior_cache = cached
    return cached[1]


def _average_irf(kind, mine, theirs, w_mine, w_theirs):
    """Exposure-weighted average of two position-independent IRFs per true-energy bin (irf/core.py ``IRFMap.stack``)."""
    from .irf import EDispKernel, EDispKernelMap, PSFMap

    w1, w2 = np.asarray(w_mine, float), np.asarray(w_theirs, float)
    total = w1 + w2
    with np.errstate(invalid="ignore", divide="ignore"):
        f1, f2 = np.nan_to_num(w1 / total), np.nan_to_num(w2 / total)
    if kind == "psf":
        if mine.rad_axis != theirs.rad_axis or mine.data.shape != theirs.data.shape:
            raise ValueError("PSF maps to stack must share their axes")
        data = mine.data * f1[:, np.newaxis] + theirs.data * f2[:, np.newaxis]
        return PSFMap(mine.energy_axis_true, mine.rad_axis, data)
    k1, k2 = mine.get_edisp_kernel(), theirs.get_edisp_kernel()
    if k1.pdf_matrix.shape != k2.pdf_matrix.shape:
        raise ValueError("edisp kernels to stack must share their axes")
    data = k1.pdf_matrix * f1[:, np.newaxis] + k2.pdf_matrix * f2[:, np.newaxis]
    return EDispKernelMap(EDispKernel([k1.axes["energy_true"], k1.axes["energy"]], data))


class _MixedEngine:
    """Joint likelihood of Cash and WStat members (``Datasets.stat_sum``, datasets/core.py:264-277: the sum of every
    dataset's own statistic, in dataset order): the Cash members share one ``LikelihoodEngine`` (one fold launch for all
    of them), every ``MapDatasetOnOff`` evaluates its npred_signal and ``wstat_sum_kernel`` through its own engine; one
    parameter vector over the union of their parameters."""

    def __init__(self, datasets):
        from .engine import LikelihoodEngine
        from .modeling.parameter import bind_mirror

        self.datasets = list(datasets)
        self._signature = LikelihoodEngine.signature(self.datasets)
        cash = [d for d in self.datasets if getattr(d, "stat_type", "cash") != "wstat"]
        self._cash = LikelihoodEngine(cash) if cash else None
        self._cash_index = {id(d): i for i, d in enumerate(cash)}
        self._onoff = {id(d): d._fit_engine() for d in self.datasets if getattr(d, "stat_type", "cash") == "wstat"}
        self.parameters, col = [], {}
        for d in self.datasets:
            for model in (d.models or []):
                for par in model.parameters:
                    if id(par) not in col:
                        col[id(par)] = len(self.parameters)
                        self.parameters.append(par)
        self.n_par = max(len(self.parameters), 1)
        self.ctx = (self._cash or next(iter(self._onoff.values()))).ctx
        self._cash_cols = np.array([col[id(p)] for p in self._cash.parameters], dtype=int) if self._cash else None
        self._onoff_cols = {k: np.array([col[id(p)] for p in e.parameters], dtype=int) for k, e in self._onoff.items()}
        self._vec = bind_mirror(self.parameters)

    def __del__(self):
        from .modeling.parameter import release_mirror

        try:
            release_mirror(self.parameters, self._vec)
        except Exception:
            pass

    def values(self):
        return self._vec.copy() if self.parameters else np.zeros(self.n_par)

    def stat_sum(self, values=None, per_dataset=False):
        if values is None:
            values = self.values()
        values = np.ascontiguousarray(values, dtype=float)
        single = values.ndim == 1
        v2 = values.reshape(1, -1) if single else values
        if v2.shape[1] != self.n_par:
            raise ValueError(f"expected {self.n_par} parameter values per vector, got {v2.shape[1]}")
        per = np.zeros((v2.shape[0], len(self.datasets)))
        if self._cash is not None:
            _, cash_per = self._cash.stat_sum(np.ascontiguousarray(v2[:, self._cash_cols]), per_dataset=True)
        for k, d in enumerate(self.datasets):
            if id(d) in self._onoff:
                per[:, k] = np.atleast_1d(self._onoff[id(d)].stat_sum(np.ascontiguousarray(v2[:, self._onoff_cols[id(d)]])))
            else:
                per[:, k] = cash_per[:, self._cash_index[id(d)]]
        stat = np.zeros(v2.shape[0])
        for k in range(len(self.datasets)):  # the reference's loop: stat_sum += dataset.stat_sum_likelihood()
            stat = stat + per[:, k]
        if per_dataset:
            return (float(stat[0]), per[0]) if single else (stat, per)
        return float(stat[0]) if single else stat


class Datasets(list):
    """Container of datasets with a joint likelihood (datasets/core.py:140)."""

    def __init__(self, datasets=None):
        if datasets is None:
            datasets = []
        elif isinstance(datasets, MapDataset):
            datasets = [datasets]
        super().__init__(datasets)
        names = [d.name for d in self]
        if len(set(names)) != len(names):
            raise ValueError("Dataset names must be unique")
        self._engine = None

    @property
    def names(self):
        return [d.name for d in self]

    @property
    def models(self):
        """Unique models of all datasets (datasets/core.py:179-199)."""
        seen, out = set(), []
        for d in self:
            for m in (d.models or []):
                if id(m) not in seen:
                    seen.add(id(m))
                    out.append(m)
        return Models(out)

    @models.setter
    def models(self, models):
        for d in self:
            d.models = models

    @property
    def parameters(self):
        return self.models.parameters.unique_parameters

    def stack_reduce(self, name=None, nan_to_num=True):
        """One dataset from all of them (datasets/core.py:530-569): the first one masked, the others stacked onto it."""
        if not len(self):
            raise ValueError("no dataset to stack")
        stacked = self[0].to_masked(name=name, nan_to_num=nan_to_num)
        for dataset in self[1:]:
            stacked.stack(dataset, nan_to_num=nan_to_num)
        return stacked

    def write(self, filename, filename_models=None, overwrite=False):
        """``Datasets.write`` (datasets/core.py:479-528): a YAML index ``{datasets: [{name, type, filename}]}``, one FITS
        file per dataset next to it and, when asked for, the models YAML (``Models.write``)."""
        import os

        import yaml

        from .modeling.io import YAML_FORMAT

        path = str(filename)
        folder = os.path.dirname(os.path.abspath(path))
        os.makedirs(folder, exist_ok=True)
        if os.path.exists(path) and not overwrite:
            raise OSError(f"File exists already: {path}")
        entries = []
        for d in self:
            entry = d.to_dict()
            d.write(os.path.join(folder, entry["filename"]), overwrite=overwrite)
            entries.append(entry)
        with open(path, "w") as fh:
            fh.write(yaml.safe_dump({"datasets": entries}, **YAML_FORMAT))
        if filename_models:
            self.models.write(filename_models, overwrite=overwrite)

    @classmethod
    def read(cls, filename, filename_models=None):
        """``Datasets.read`` (datasets/core.py:435-477): the datasets of a YAML index (file names relative to it) and,
        when given, their models (``datasets.models = Models.read(...)``: every dataset picks the models that name it
        or name no dataset)."""
        import os

        import yaml

        path = str(filename)
        with open(path) as fh:
            index = yaml.safe_load(fh.read())
        folder = os.path.dirname(os.path.abspath(path))
        kinds = {MapDataset.tag: MapDataset, MapDatasetOnOff.tag: MapDatasetOnOff}
        datasets = []
        for entry in index["datasets"]:
            if entry["type"] not in kinds:
                raise NotImplementedError(f"dataset type {entry['type']!r} is outside the accelerated path "
                                          f"(available: {sorted(kinds)})")
            fname = entry["filename"]
            if os.path.exists(os.path.join(folder, fname)):
                fname = os.path.join(folder, fname)
            datasets.append(kinds[entry["type"]].read(fname, name=entry["name"]))
        out = cls(datasets)
        if filename_models:
            out.models = Models.read(filename_models)
        return out

    def _get_engine(self):
        from .engine import LikelihoodEngine

        if any(getattr(d, "stat_type", "cash") == "wstat" for d in self):
            # ON/OFF members bring their own statistic (WStat): Cash members in one engine, the others one by one
            stale = (self._engine is None or not isinstance(self._engine, _MixedEngine)
                     or self._engine._signature != LikelihoodEngine.signature(self))
            if not stale:
                stale = any(d._device is None or d._device._uploaded != d._data_fingerprint()
                            for d in self if getattr(d, "stat_type", "cash") != "wstat")
            if stale:
                self._engine = _MixedEngine(self)
            return self._engine
        stale = self._engine is None or self._engine._signature != LikelihoodEngine.signature(self)
        if not stale:
            stale = any(d._device is None or d._device._uploaded != d._data_fingerprint() for d in self)
        if stale:
            self._engine = LikelihoodEngine(self)
        return self._engine

    def stat_sum(self):
        """Joint fit statistic (datasets/core.py:264-277): the likelihood of all datasets (GPU) plus the priors of
        the parameters that carry one (host)."""
        eng = self._get_engine()
        value = eng.stat_sum()
        priors = _parameters_with_prior(eng)
        if priors:
            value += sum(p.prior(p) for p in priors)
        return value

    def stat_sum_likelihood(self):
        """Total statistic without the priors (datasets/core.py:279-284)."""
        return self._get_engine().stat_sum()

    def stat_sum_batch(self, values, per_dataset=False):
        """Joint statistic for a batch of full parameter vectors [B, n_par] in ONE launch
        (what ``Fit.stat_profile`` / ``stat_surface`` loop over, modeling/fit.py:351-519).  Like ``stat_sum`` it
        includes the parameters' priors (datasets/core.py:264-277), evaluated on the host per vector; with
        ``per_dataset=True`` the likelihood terms alone are returned, one column per dataset."""
        eng = self._get_engine()
        out = eng.stat_sum(values, per_dataset=per_dataset)
        if not per_dataset and any(p.prior is not None for p in eng.parameters):
            vals = np.atleast_2d(np.asarray(values, dtype=float))
            prior = np.zeros(vals.shape[0])
            for col, par in enumerate(eng.parameters):
                if par.prior is not None:
                    prior += par.prior.evaluate_many(vals[:, col])
            out = out + (float(prior[0]) if np.ndim(out) == 0 else prior)
        return out

    def parameter_vector(self):
        eng = self._get_engine()
        return eng.values(), eng.parameters

    def __getitem__(self, key):
        if isinstance(key, str):
            for d in self:
                if d.name == key:
                    return d
            raise KeyError(key)
        if isinstance(key, slice):
            return Datasets(list.__getitem__(self, key))
        return list.__getitem__(self, key)

    def __repr__(self):
        return f"Datasets({self.names})"
