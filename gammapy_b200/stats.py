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


def wstat_sum_cuda(n_on, n_off, alpha, mu_sig, mask=None, return_array=False):
    """``WStatFitStatistic.stat_sum_dataset`` (stats/fit_statistics.py:332-361): sum over the masked bins of
    ``np.nan_to_num(wstat(n_on, n_off, alpha, mu_sig))`` (:142-252) on the GPU.  Arrays of one shape: numpy (copied to
    the device as float64) or CUDA torch tensors (ON / OFF / alpha float32 or float64, ``mu_sig`` float64, ``mask``
    uint8 / bool).  ``return_array``: also the per-bin statistic (``stat_array_dataset``) as a numpy array."""
    import torch

    ctx = _lib.Context.get()
    dev = ctx.torch_device

    def to_dev(a, dtype=None):
        if isinstance(a, torch.Tensor):
            t = a.to(dev)
            return t.contiguous() if dtype is None else t.to(dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64 if dtype is None else None)).to(dev) \
            if dtype is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dtype).contiguous()

    t_on, t_off, t_al = to_dev(n_on), to_dev(n_off), to_dev(alpha)
    kinds = {t.dtype for t in (t_on, t_off, t_al)}
    if kinds == {torch.float32}:
        f64 = 0
    else:
        t_on, t_off, t_al = (t.to(torch.float64).contiguous() for t in (t_on, t_off, t_al))
        f64 = 1
    t_mu = to_dev(mu_sig, torch.float64)
    if not (t_on.shape == t_off.shape == t_al.shape == t_mu.shape):
        raise ValueError("n_on, n_off, alpha and mu_sig must have the same shape")
    t_mask = None if mask is None else to_dev(mask, torch.uint8)
    if t_mask is not None and t_mask.shape != t_on.shape:
        raise ValueError("mask must have the shape of the counts")
    per_bin = torch.empty(t_on.shape, dtype=torch.float64, device=dev) if return_array else None
    torch.cuda.current_stream(dev).synchronize()  # the library runs on its own stream
    out = C.c_double(0.0)
    _lib.check(ctx.lib.gpy_wstat_sum_device(
        ctx.handle, C.c_void_p(t_on.data_ptr()), C.c_void_p(t_off.data_ptr()), C.c_void_p(t_al.data_ptr()), f64,
        C.c_void_p(t_mu.data_ptr()), C.c_void_p(t_mask.data_ptr()) if t_mask is not None else None, t_on.numel(),
        C.c_void_p(per_bin.data_ptr()) if per_bin is not None else None, C.byref(out)))
    return (out.value, per_bin.cpu().numpy()) if return_array else out.value


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


def cash(n_on, mu_on, truncation_value=1e-25):
    """Per-bin Cash statistic 2 (mu - n log mu) with mu truncated from below (stats/fit_statistics.py:29-79); host numpy."""
    n_on, mu_on = np.asarray(n_on, dtype=float), np.asarray(mu_on, dtype=float)
    mu_on = np.where(mu_on <= truncation_value, truncation_value, mu_on)
    with np.errstate(divide="ignore", invalid="ignore"):
        return 2 * (mu_on - n_on * np.log(mu_on))


def cash_sqrt_ts(n_on, mu_bkg):
    """``CashCountsStatistic(n_on, mu_bkg).sqrt_ts`` (stats/counts_statistic.py:44-58, 250-288): the significance of the
    excess n_on - mu_bkg for a known background, sign of the excess times sqrt(stat_null - stat_max)."""
    n_on, mu_bkg = np.asarray(n_on, dtype=float), np.asarray(mu_bkg, dtype=float)
    ts = np.clip(cash(n_on, mu_bkg + 0) - cash(n_on, n_on), 0, None)
    return np.sign(n_on - mu_bkg) * np.sqrt(ts)


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
