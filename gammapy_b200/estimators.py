"""TS-map inner loop on the GPU (estimators/map/ts.py in the reference).

``TSMapEstimator.estimate_flux_map`` (ts.py:516-591) distributes ``_ts_value`` (:1098-1153) over processes: per
valid pixel a ``SimpleMapDataset.from_arrays`` cutout and ``BrentqFluxEstimator.run`` (:1053-1095), i.e. norm
bounds, a Brent root of the Cash derivative, the error from the second derivative and TS = stat(0) - stat(norm).
Here all pixels are one launch (one warp per pixel) through ``gpy_ts_map``.

Scope: the default selection (no "ul" / "errn-errp" / "stat_scan" / "sensitivity"), ``ts_threshold=None``.
Preparing the input maps from a ``MapDataset`` (``estimate_fit_input_maps``: exposure in reco energy, kernel
estimation, default mask) is the caller's job: pass the arrays that method returns.
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["RESULT_NAMES", "compute_ts_map", "BrentqFluxEstimator", "TSMapEstimator", "ParameterEstimator", "FluxPointsEstimator"]

RESULT_NAMES = ("norm", "norm_err", "niter", "ts", "stat", "stat_null", "success", "npred", "npred_excess")


def _cube(a, name):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 2:
        a = a[np.newaxis]
    if a.ndim != 3:
        raise ValueError(f"{name} must have shape (n_planes, ny, nx)")
    return a


def compute_ts_map(counts, background, exposure, kernel, mask=None, rtol=0.01, max_niter=100, n_sigma=1.0):
    """One-parameter norm fit at every pixel of ``mask``.

    ``counts`` / ``background`` / ``exposure``: (n_planes, ny, nx); several datasets of one image geometry are
    concatenated along the first axis (as ``_ts_value`` concatenates their cutouts); ``kernel``: (n_planes, ky, kx)
    with odd sides.  Returns a dict of (ny, nx) float64 arrays named as the reference's result dict; positions
    outside ``mask`` hold NaN (``Map.from_geom(..., data=np.nan)``, ts.py:589)."""
    counts, background = _cube(counts, "counts"), _cube(background, "background")
    exposure, kernel = _cube(exposure, "exposure"), _cube(kernel, "kernel")
    if not (counts.shape == background.shape == exposure.shape):
        raise ValueError("counts, background and exposure must have the same shape")
    if kernel.shape[0] != counts.shape[0]:
        raise ValueError("kernel needs one plane per plane of the maps")
    n_planes, ny, nx = counts.shape
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.asarray(mask).reshape(ny, nx) != 0, dtype=np.uint8)
    out = np.empty((len(RESULT_NAMES), ny, nx), dtype=np.float64)
    ctx = _lib.Context.get()

    def ptr(a):
        return a.ctypes.data_as(C.c_void_p)

    _lib.check(ctx.lib.gpy_ts_map(ctx.handle, ptr(counts), ptr(background), ptr(exposure), n_planes, ny, nx,
                                  ptr(kernel), kernel.shape[1], kernel.shape[2], ptr(m) if m is not None else None,
                                  float(rtol), int(max_niter), float(n_sigma), ptr(out)))
    result = {name: out[k] for k, name in enumerate(RESULT_NAMES)}
    result["success"] = np.where(np.isnan(out[6]), False, out[6] != 0)
    return result


class BrentqFluxEstimator:
    """Single-parameter flux estimator (ts.py:787-818): here only a parameter holder for ``TSMapEstimator``."""

    tag = "BrentqFluxEstimator"

    def __init__(self, rtol, n_sigma, n_sigma_ul=2, n_sigma_sensitivity=None, selection_optional=None, max_niter=100,
                 ts_threshold=None, norm=None):
        if selection_optional:
            raise NotImplementedError("only the default selection is on the GPU path (no ul / errn-errp / stat_scan)")
        if ts_threshold is not None:
            raise NotImplementedError("ts_threshold is not on the GPU path")
        self.rtol, self.n_sigma, self.n_sigma_ul, self.max_niter = rtol, n_sigma, n_sigma_ul, max_niter


class TSMapEstimator:
    """The fit stage of the reference's ``TSMapEstimator`` (ts.py:68-300 for the parameters).

    ``estimate_flux_map(maps)`` takes what ``estimate_fit_input_maps`` (:444-498) returns for each dataset --
    dicts with ``counts``, ``background``, ``exposure``, ``kernel`` (arrays (n_e, ny, nx) / (n_e, ky, kx)) and
    ``mask`` -- and returns the result maps of ``estimate_flux_map`` (:516-591) as arrays, plus ``sqrt_ts``."""

    tag = "TSMapEstimator"

    def __init__(self, n_sigma=1, n_sigma_ul=2, rtol=0.01, max_niter=100, selection_optional=None):
        self._flux_estimator = BrentqFluxEstimator(rtol=rtol, n_sigma=n_sigma, n_sigma_ul=n_sigma_ul,
                                                   selection_optional=selection_optional, max_niter=max_niter)

    @staticmethod
    def _estimate_mask_2d(map_dict):
        """ts.py:500-505: any energy plane selected by the mask and a positive summed background."""
        mask = np.asarray(map_dict["mask"]).astype(bool)
        background = np.asarray(map_dict["background"], dtype=float)
        if mask.ndim == 3:
            mask = np.logical_or.reduce(mask, axis=0)
        if background.ndim == 3:
            background = background.sum(axis=0)
        return mask & (background > 0)

    def estimate_flux_map(self, maps):
        if isinstance(maps, dict):
            maps = [maps]
        mask_2d = np.logical_or.reduce([self._estimate_mask_2d(m) for m in maps])
        if not np.any(mask_2d):
            raise ValueError("No valid positions found. Check that the dataset background is defined and not only "
                             "zeros, or that the mask is not all False.")
        stack = {k: np.concatenate([_cube(m[k], k) for m in maps], axis=0)
                 for k in ("counts", "background", "exposure", "kernel")}
        fe = self._flux_estimator
        result = compute_ts_map(stack["counts"], stack["background"], stack["exposure"], stack["kernel"], mask=mask_2d,
                                rtol=fe.rtol, max_niter=fe.max_niter, n_sigma=fe.n_sigma)
        with np.errstate(invalid="ignore"):
            ts = result["ts"]
            result["sqrt_ts"] = np.where(ts > 0, np.sqrt(np.abs(ts)), -np.sqrt(np.abs(ts)))
        return result


# -------------------------------------------------------------------------------------------------------------
# ParameterEstimator (estimators/parameter.py:20-440): best fit, delta(TS) against a null value, asymmetric
# errors, upper limit and profile of ONE parameter, on top of Fit (whose gradients, scans and confidence searches
# are batched GPU launches).  Works for Datasets and ShardedDatasets alike.
# -------------------------------------------------------------------------------------------------------------
class _restore_status:
    """parameters.restore_status(): values and frozen flags come back on exit (modeling/parameter.py)."""

    def __init__(self, parameters):
        self.parameters = list(parameters)

    def __enter__(self):
        self.saved = [(p, p.value, p.frozen) for p in self.parameters]
        return self

    def __exit__(self, *exc):
        for p, value, frozen in self.saved:
            p.value, p.frozen = value, frozen


class ParameterEstimator:
    tag = "ParameterEstimator"
    _available_selection_optional = ["errn-errp", "ul", "scan"]

    def __init__(self, n_sigma=1, n_sigma_ul=2, null_value=1e-150, selection_optional=None, fit=None, reoptimize=True):
        from .modeling import Fit

        self.n_sigma, self.n_sigma_ul, self.null_value = n_sigma, n_sigma_ul, null_value
        if selection_optional == "all":
            selection_optional = list(self._available_selection_optional)
        self.selection_optional = list(selection_optional or [])
        unknown = set(self.selection_optional) - set(self._available_selection_optional)
        if unknown:
            raise NotImplementedError(f"selection_optional {sorted(unknown)} is not available here")
        self.fit = fit if fit is not None else Fit()
        self.reoptimize = reoptimize

    @staticmethod
    def _parameter(datasets, parameter):
        if isinstance(parameter, str):
            return datasets.parameters[parameter]
        return parameter

    def estimate_best_fit(self, datasets, parameter):
        """parameter.py:108-141."""
        result = self.fit.run(datasets)
        return {f"{parameter.name}": parameter.value, "stat": result.optimize_result.total_stat,
                "success": result.success, f"{parameter.name}_err": parameter.error * self.n_sigma}

    def estimate_ts(self, datasets, parameter):
        """parameter.py:143-181: stat at the null value (other parameters re-fitted when ``reoptimize``) minus the
        best-fit stat."""
        stat = datasets.stat_sum()
        with _restore_status(datasets.parameters):
            parameter.value = self.null_value
            if self.reoptimize:
                parameter.frozen = True
                if any(not p.frozen for p in datasets.parameters):
                    self.fit.optimize(datasets)
            stat_null = datasets.stat_sum()
        return {"ts": stat_null - stat, "stat_null": stat_null}

    def estimate_errn_errp(self, datasets, parameter):
        """parameter.py:183-219."""
        self.fit.optimize(datasets)
        res = self.fit.confidence(datasets, parameter, sigma=self.n_sigma, reoptimize=self.reoptimize)
        return {f"{parameter.name}_errp": res["errp"], f"{parameter.name}_errn": res["errn"]}

    def estimate_ul(self, datasets, parameter):
        """parameter.py:258-286."""
        self.fit.optimize(datasets)
        res = self.fit.confidence(datasets, parameter, sigma=self.n_sigma_ul, reoptimize=self.reoptimize)
        return {f"{parameter.name}_ul": res["errp"] + parameter.value}

    def estimate_scan(self, datasets, parameter):
        """parameter.py:221-256."""
        self.fit.optimize(datasets)
        profile = self.fit.stat_profile(datasets, parameter, reoptimize=self.reoptimize)
        return {f"{parameter.name}_scan": np.asarray(parameter.scan_values, dtype=float), "stat_scan": profile["stat_scan"]}

    def run(self, datasets, parameter):
        """parameter.py:354-411 (without the per-dataset counts / npred bookkeeping)."""
        parameter = self._parameter(datasets, parameter)
        with _restore_status(datasets.parameters):
            if not self.reoptimize:
                for p in datasets.parameters:
                    p.frozen = True
                parameter.frozen = False
            result = self.estimate_best_fit(datasets, parameter)
            result.update(self.estimate_ts(datasets, parameter))
            if "errn-errp" in self.selection_optional:
                result.update(self.estimate_errn_errp(datasets, parameter))
            if "ul" in self.selection_optional:
                result.update(self.estimate_ul(datasets, parameter))
            if "scan" in self.selection_optional:
                result.update(self.estimate_scan(datasets, parameter))
        return result


# -------------------------------------------------------------------------------------------------------------
# FluxPointsEstimator (estimators/points/sed.py:41-300, estimators/flux.py): one norm fit of the source per energy
# group.  The reference slices the datasets to the group (Datasets.slice_by_energy) and scales the source's spectrum
# with a norm parameter (ScaleSpectralModel); here the fit mask is restricted to the group's reco bins -- the same
# likelihood terms, on the resident cubes -- and the norm is the source's amplitude in units of its reference value
# (the amplitude is a pure scale factor of all three spectral models).  Every fit is ParameterEstimator on the GPU
# likelihood.
# -------------------------------------------------------------------------------------------------------------
class FluxPointsEstimator:
    tag = "FluxPointsEstimator"

    def __init__(self, energy_edges, source=0, n_sigma=1, n_sigma_ul=2, selection_optional=None, fit=None, reoptimize=False):
        self.energy_edges = np.asarray(energy_edges, dtype=float)
        if self.energy_edges.ndim != 1 or len(self.energy_edges) < 2 or np.any(np.diff(self.energy_edges) <= 0):
            raise ValueError("energy_edges must be an increasing 1-D array of at least two values [TeV]")
        self.source = source
        self.reoptimize = reoptimize
        self._estimator = ParameterEstimator(n_sigma=n_sigma, n_sigma_ul=n_sigma_ul, null_value=0.0,
                                             selection_optional=selection_optional, fit=fit, reoptimize=True)

    def _source_model(self, datasets):
        from .modeling.models import SkyModel

        sky = [m for ds in datasets for m in (ds.models or []) if isinstance(m, SkyModel)]
        unique = list({id(m): m for m in sky}.values())
        if isinstance(self.source, str):
            found = [m for m in unique if m.name == self.source]
            if not found:
                raise KeyError(self.source)
            return found[0]
        return unique[self.source]

    @staticmethod
    def _group(axis, emin, emax):
        """Reco bins of ``axis`` inside [emin, emax] (the bins Datasets.slice_by_energy keeps: edges matched to 1e-3)."""
        edges = axis.edges
        lo = np.isclose(edges[:-1], emin, rtol=1e-3) | (edges[:-1] > emin)
        hi = np.isclose(edges[1:], emax, rtol=1e-3) | (edges[1:] < emax)
        return lo & hi

    def run(self, datasets):
        from .datasets import Datasets, MapDataset
        from .maps import Map

        if isinstance(datasets, MapDataset):
            datasets = Datasets([datasets])
        model = self._source_model(datasets)
        amplitude = model.spectral_model.amplitude
        amp_ref = amplitude.value
        if not amp_ref > 0:
            raise ValueError("the source needs a positive reference amplitude")
        saved_masks = [ds.mask_fit for ds in datasets]
        rows = []
        try:
            for emin, emax in zip(self.energy_edges[:-1], self.energy_edges[1:]):
                groups = [self._group(ds._geom.axes["energy"], emin, emax) for ds in datasets]
                contributing = [g for g in groups if np.any(g)]
                e_lo = min(ds._geom.axes["energy"].edges[:-1][g].min() for ds, g in zip(datasets, groups) if np.any(g)) \
                    if contributing else emin
                e_hi = max(ds._geom.axes["energy"].edges[1:][g].max() for ds, g in zip(datasets, groups) if np.any(g)) \
                    if contributing else emax
                e_ref = float(np.sqrt(e_lo * e_hi))
                ref_dnde = float(model.spectral_model(e_ref))
                row = {"e_ref": e_ref, "e_min": float(e_lo), "e_max": float(e_hi), "ref_dnde": ref_dnde}
                if not contributing:
                    row.update({"norm": np.nan, "norm_err": np.nan, "ts": np.nan, "stat": np.nan, "stat_null": np.nan,
                                "success": False})
                    rows.append(row)
                    continue
                for ds, g, base in zip(datasets, groups, saved_masks):
                    sel = np.zeros(ds._geom.data_shape, dtype=bool)
                    sel[g] = True
                    if base is not None:
                        sel &= np.asarray(base.data, dtype=bool)
                    ds.mask_fit = Map.from_geom(ds._geom, data=sel, dtype=bool)
                pars = list(datasets.parameters)
                with _restore_status(pars):
                    for p in pars:  # the spectral shape of the source is fixed; other parameters only with reoptimize
                        if p is not amplitude and (not self.reoptimize or any(p is q for q in model.parameters)):
                            p.frozen = True
                    amplitude.frozen = False
                    res = self._estimator.run(datasets, amplitude)
                row.update({"norm": res["amplitude"] / amp_ref, "norm_err": res["amplitude_err"] / amp_ref, "ts": res["ts"],
                            "stat": res["stat"], "stat_null": res["stat_null"], "success": bool(res["success"])})
                for key in ("amplitude_errn", "amplitude_errp", "amplitude_ul"):
                    if key in res:
                        row[key.replace("amplitude", "norm")] = res[key] / amp_ref
                rows.append(row)
        finally:
            for ds, base in zip(datasets, saved_masks):
                ds.mask_fit = base
            amplitude.value = amp_ref
        out = {k: np.array([r.get(k, np.nan) for r in rows]) for k in rows[0]}
        for r in rows:
            for k in r:
                if k not in out:
                    out[k] = np.array([q.get(k, np.nan) for q in rows])
        with np.errstate(invalid="ignore"):
            out["sqrt_ts"] = np.where(out["ts"] > 0, np.sqrt(np.abs(out["ts"])), -np.sqrt(np.abs(out["ts"])))
        out["dnde"] = out["norm"] * out["ref_dnde"]
        out["dnde_err"] = out["norm_err"] * out["ref_dnde"]
        for key in ("norm_errn", "norm_errp", "norm_ul"):
            if key in out:
                out[key.replace("norm", "dnde")] = out[key] * out["ref_dnde"]
        return out
