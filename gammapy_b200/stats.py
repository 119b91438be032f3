"""Compiled fit-statistics registry with a ``cuda`` backend (utils/compilation.py:35-87 in the reference).

``get_fit_statistics_compiled()`` returns the same dict keys the reference's cython / jit backends
provide for the Cash path; the callables take caller-owned 1-D float64 host arrays (already masked)
and return a float, exactly like ``cash_sum_cython`` (stats/fit_statistics_cython.pyx:17-85).
This is the narrow operator plug-in (boundary B2); the fused path uses ``Datasets.stat_sum``.
"""
import ctypes as C

import numpy as np

from . import _lib

TRUNCATION_VALUE = 1e-25
COMPILATION_BACKEND_DEFAULT = "cuda"


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64).ravel()


def cash_sum_cuda(counts, npred):
    counts, npred = _as_f64(counts), _as_f64(npred)
    if counts.shape != npred.shape:
        raise ValueError("counts and npred must have the same length")
    ctx = _lib.Context.get()
    out = C.c_double(0.0)
    _lib.check(ctx.lib.gpy_cash_sum(ctx.handle, counts.ctypes.data_as(C.c_void_p), npred.ctypes.data_as(C.c_void_p),
                                    counts.size, C.byref(out)))
    return out.value


def weighted_cash_sum_cuda(counts, npred, weight):
    counts, npred, weight = _as_f64(counts), _as_f64(npred), _as_f64(weight)
    if not (counts.shape == npred.shape == weight.shape):
        raise ValueError("counts, npred and weight must have the same length")
    ctx = _lib.Context.get()
    out = C.c_double(0.0)
    _lib.check(ctx.lib.gpy_weighted_cash_sum(ctx.handle, counts.ctypes.data_as(C.c_void_p),
                                             npred.ctypes.data_as(C.c_void_p), weight.ctypes.data_as(C.c_void_p),
                                             counts.size, C.byref(out)))
    return out.value


def f_cash_root_cuda(x, counts, background, model):
    """f_cash_root_cython (stats/fit_statistics_cython.pyx:90-121): derivative of the Cash statistic wrt the norm."""
    counts, background, model = _as_f64(counts), _as_f64(background), _as_f64(model)
    if not (counts.shape == background.shape == model.shape):
        raise ValueError("counts, background and model must have the same length")
    ctx = _lib.Context.get()
    out = C.c_double(0.0)
    _lib.check(ctx.lib.gpy_f_cash_root(ctx.handle, float(x), counts.ctypes.data_as(C.c_void_p),
                                       background.ctypes.data_as(C.c_void_p), model.ctypes.data_as(C.c_void_p),
                                       counts.size, C.byref(out)))
    return out.value


def norm_bounds_cuda(counts, background, model):
    """norm_bounds_cython (.pyx:124-157): (b_min, b_max, -sn_min_total)."""
    counts, background, model = _as_f64(counts), _as_f64(background), _as_f64(model)
    if not (counts.shape == background.shape == model.shape):
        raise ValueError("counts, background and model must have the same length")
    ctx = _lib.Context.get()
    out = (C.c_double * 3)()
    _lib.check(ctx.lib.gpy_norm_bounds(ctx.handle, counts.ctypes.data_as(C.c_void_p),
                                       background.ctypes.data_as(C.c_void_p), model.ctypes.data_as(C.c_void_p),
                                       counts.size, out))
    return out[0], out[1], out[2]


def get_fit_statistics_compiled(backend=None):
    backend = backend or COMPILATION_BACKEND_DEFAULT
    if backend != "cuda":
        raise ValueError(f"Unknown compilation backend {backend!r}: this package only ships 'cuda'")
    return {
        "TRUNCATION_VALUE": TRUNCATION_VALUE,
        "cash_sum_compiled": cash_sum_cuda,
        "weighted_cash_sum_compiled": weighted_cash_sum_cuda,
        "f_cash_root_compiled": f_cash_root_cuda,
        "norm_bounds_compiled": norm_bounds_cuda,
    }
