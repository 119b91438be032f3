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

__all__ = ["RESULT_NAMES", "compute_ts_map", "BrentqFluxEstimator", "TSMapEstimator", "ParameterEstimator"]

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
