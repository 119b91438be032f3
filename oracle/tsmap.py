"""ORACLE (test infrastructure, never on the product path): CPU restatement of the TS-map inner loop.

Follows, relative to /root/reference/gammapy:
  estimators/map/ts.py:43-65     _extract_array
  estimators/map/ts.py:703-783   SimpleMapDataset (from_arrays, stat_derivative, stat_2nd_derivative, npred)
  estimators/map/ts.py:819-870   BrentqFluxEstimator.estimate_best_fit
  estimators/map/ts.py:1053-1095 BrentqFluxEstimator.run (npred, npred_excess, stat)
  estimators/map/ts.py:1098-1153 _ts_value (datasets concatenated)
  stats/fit_statistics_cython.pyx:55-157  cash_sum / f_cash_root / norm_bounds

The three compiled statistics come from the reference's own Cython file compiled into ``oracle/_ref`` when it is
available (``use_reference=True``), else from the loops below.  The root finder is ``scipy.optimize.brentq`` -- a
third-party dependency of the reference (scipy >= 1.13, < 1.18; installed here: see ``scipy.__version__``), not
under /root/reference -- when scipy is importable, else :func:`brentq`, a restatement of its published C code
(scipy/optimize/Zeros/brentq.c) that ``tests/test_oracle_tsmap.py`` pins against the installed scipy
(root, iterations and function calls equal on every case).
"""
import math

import numpy as np

TRUNCATION_VALUE = 1e-25
RESULT_NAMES = ("norm", "norm_err", "niter", "ts", "stat", "stat_null", "success", "npred", "npred_excess")


# ---- stats/fit_statistics_cython.pyx:90-157 ------------------------------------------------------------------
def f_cash_root(x, counts, background, model):
    total = 0.0
    with np.errstate(divide="ignore", invalid="ignore"):  # cdivision: x / 0 is inf, not an exception
        for c, b, m in zip(np.asarray(counts, dtype=float), np.asarray(background, dtype=float),
                           np.asarray(model, dtype=float)):
            if m > 0:
                if c > 0:
                    total += m * (1 - c / (x * m + b))
                else:
                    total += m
    return float(2 * total)


def norm_bounds(counts, background, model):
    s_model = s_counts = 0.0
    sn_min = sn_min_total = 1e14
    c_min = 1.0
    for c, b, m in zip(counts, background, model):
        if c > 0:
            s_counts += c
            if m > 0:
                sn = b / m
                if sn < sn_min:
                    sn_min, c_min = sn, c
        if m > 0:
            s_model += m
            sn = b / m
            if sn < sn_min_total:
                sn_min_total = sn
    with np.errstate(divide="ignore", invalid="ignore"):
        b_min = np.float64(c_min) / np.float64(s_model) - sn_min
        b_max = np.float64(s_counts) / np.float64(s_model) - sn_min
    return float(b_min), float(b_max), -sn_min_total


def cash_sum(counts, npred):
    total = 0.0
    log_trunc = math.log(float(np.float32(TRUNCATION_VALUE)))  # the Cython constant is narrowed to C float
    for c, mu in zip(counts, npred):
        if mu > TRUNCATION_VALUE:
            npr = mu
            lognpr = math.log(mu) if c > 0 else 0.0
        else:
            npr = TRUNCATION_VALUE
            lognpr = log_trunc
        total += npr
        if c > 0:
            total -= c * lognpr
    return 2 * total


def _compiled(use_reference):
    if use_reference:
        from . import cash

        ref = cash.reference_module()
        if ref is not None:
            return ref.cash_sum_cython, ref.f_cash_root_cython, ref.norm_bounds_cython
    return cash_sum, f_cash_root, norm_bounds


# ---- scipy/optimize/Zeros/brentq.c ----------------------------------------------------------------------------
class BrentError(Exception):
    """scipy raises ValueError (sign / nan bounds) or RuntimeError (not converged, nan value): both are caught by
    the reference (ts.py:855)."""


def brentq(f, xa, xb, xtol=2e-12, rtol=8.881784197001252e-16, maxiter=100):
    """Returns (root, iterations, function_calls); raises BrentError where scipy raises."""
    if xtol <= 0 or rtol < 4 * np.finfo(float).eps / 1.0 or maxiter < 1:
        raise BrentError("bad tolerance or maxiter")
    if math.isnan(xa) or math.isnan(xb):
        raise BrentError("nan bounds")
    xpre, xcur = xa, xb
    xblk = fblk = spre = scur = 0.0
    fpre, fcur = f(xpre), f(xcur)
    calls = 2
    if math.isnan(fpre) or math.isnan(fcur):
        raise BrentError("nan function value")
    if fpre == 0:
        return xpre, 0, calls
    if fcur == 0:
        return xcur, 0, calls
    if math.copysign(1.0, fpre) == math.copysign(1.0, fcur):
        raise BrentError("f(a) and f(b) must have different signs")
    for it in range(1, maxiter + 1):
        if fpre != 0 and fcur != 0 and math.copysign(1.0, fpre) != math.copysign(1.0, fcur):
            xblk, fblk = xpre, fpre
            spre = scur = xcur - xpre
        if abs(fblk) < abs(fcur):
            xpre, xcur, xblk = xcur, xblk, xcur
            fpre, fcur, fblk = fcur, fblk, fcur
        delta = (xtol + rtol * abs(xcur)) / 2
        sbis = (xblk - xcur) / 2
        if fcur == 0 or abs(sbis) < delta:
            return xcur, it, calls
        if abs(spre) > delta and abs(fcur) < abs(fpre):
            if xpre == xblk:
                stry = -fcur * (xcur - xpre) / (fcur - fpre)
            else:
                dpre = (fpre - fcur) / (xpre - xcur)
                dblk = (fblk - fcur) / (xblk - xcur)
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre))
            if 2 * abs(stry) < min(abs(spre), 3 * abs(sbis) - delta):
                spre, scur = scur, stry
            else:
                spre = scur = sbis
        else:
            spre = scur = sbis
        xpre, fpre = xcur, fcur
        if abs(scur) > delta:
            xcur += scur
        else:
            xcur += delta if sbis > 0 else -delta
        fcur = f(xcur)
        calls += 1
        if math.isnan(fcur):
            raise BrentError("nan function value")
    raise BrentError("failed to converge")


def _root(f, a, b, rtol, maxiter, use_scipy):
    if use_scipy:
        try:
            import scipy.optimize
        except ImportError:
            use_scipy = False
    if use_scipy:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                root, info = scipy.optimize.brentq(f=f, a=a, b=b, maxiter=maxiter, full_output=True, rtol=rtol)
                return root, info.iterations, info.converged
            except (RuntimeError, ValueError):
                return None
    try:
        root, iterations, _ = brentq(f, a, b, rtol=rtol, maxiter=maxiter)
        return root, iterations, True
    except BrentError:
        return None


# ---- estimators/map/ts.py ---------------------------------------------------------------------------------------
def extract_array(array, shape, position):
    """ts.py:43-65, with what falls outside the array absent (the reference pads the maps with zeros first, and
    all-zero bins are removed by from_arrays, :777)."""
    y_w, x_w = shape[1] // 2, shape[2] // 2
    y_lo, y_hi = position[0] - y_w, position[0] + y_w + 1
    x_lo, x_hi = position[1] - x_w, position[1] + x_w + 1
    out = np.zeros((array.shape[0], shape[1], shape[2]), dtype=array.dtype)
    ys, xs = max(y_lo, 0), max(x_lo, 0)
    ye, xe = min(y_hi, array.shape[1]), min(x_hi, array.shape[2])
    if ye > ys and xe > xs:
        out[:, ys - y_lo:ye - y_lo, xs - x_lo:xe - x_lo] = array[:, ys:ye, xs:xe]
    return out


def ts_value(position, counts, background, exposure, kernel, rtol=0.01, max_niter=100, n_sigma=1.0,
             use_reference=True, use_scipy=True):
    """_ts_value + BrentqFluxEstimator.run for one pixel; returns the result dict of the default selection."""
    cash_c, froot_c, bounds_c = _compiled(use_reference)
    c = extract_array(counts, kernel.shape, position)
    b = extract_array(background, kernel.shape, position)
    e = extract_array(exposure, kernel.shape, position)
    model = kernel * e
    invalid = (c == 0) & (b == 0) & (model == 0)
    c, b, model = (np.ascontiguousarray(a[~invalid], dtype=float) for a in (c, b, model))

    norm_min, norm_max, norm_min_total = bounds_c(c, b, model)
    if c.sum() <= 0 or model.sum() <= 0:
        norm, niter, success = norm_min_total, 0, True
    else:
        found = _root(lambda x: froot_c(x, c, b, model), norm_min, norm_max, rtol, max_niter, use_scipy)
        if found is None:
            norm, niter, success = norm_min_total, max_niter, False
        else:
            norm, niter, success = max(found[0], norm_min_total), found[1], found[2]
    with np.errstate(invalid="ignore", divide="ignore"):
        top = model ** 2 * c
        bottom = (b + norm * model) ** 2
        d2 = (top / bottom)[~(bottom == 0)].sum()
        norm_err = np.sqrt(1 / d2) * n_sigma
    stat = cash_c(c, b + norm * model)
    stat_null = cash_c(c, b + 0 * model)
    npred = (b + norm * model).sum()
    return {"norm": norm, "norm_err": float(norm_err), "niter": niter, "ts": stat_null - stat, "stat": stat,
            "stat_null": stat_null, "success": success, "npred": float(npred),
            "npred_excess": float(npred - (b + 0 * model).sum())}


def ts_map(counts, background, exposure, kernel, mask=None, **kwargs):
    """All pixels of ``mask``; returns {name: (ny, nx) float array}, NaN where not fitted."""
    counts, background, exposure, kernel = (np.asarray(a, dtype=float) for a in (counts, background, exposure, kernel))
    _, ny, nx = counts.shape
    out = {name: np.full((ny, nx), np.nan) for name in RESULT_NAMES}
    for j in range(ny):
        for i in range(nx):
            if mask is not None and not mask[j, i]:
                continue
            res = ts_value((j, i), counts, background, exposure, kernel, **kwargs)
            for name in RESULT_NAMES:
                out[name][j, i] = res[name]
    return out
