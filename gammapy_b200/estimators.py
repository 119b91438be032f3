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

__all__ = ["RESULT_NAMES", "compute_ts_map", "BrentqFluxEstimator", "TSMapEstimator"]

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
