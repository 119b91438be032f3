"""Oracle: Cash fit statistic (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows (paths relative to /root/reference/gammapy):

* ``cash``                       stats/fit_statistics.py:29-79   (per-bin array form, no ``n > 0`` guard)
* ``cash_sum_cython``            stats/fit_statistics_cython.pyx:55-85
* ``weighted_cash_sum_cython``   stats/fit_statistics_cython.pyx:17-51
* ``CashFitStatistic.stat_sum_dataset``          stats/fit_statistics.py:281-293
* ``WeightedCashFitStatistic.stat_sum_dataset``  stats/fit_statistics.py:304-321

``cash_sum`` / ``weighted_cash_sum`` restate the serial Cython loops with numpy (the only difference is
the fp64 summation order).  ``reference_module()`` returns the reference's Cython module compiled
unmodified by ``oracle/build_ref.py`` when it is available – that one is bit-for-bit the reference.
"""
import numpy as np

TRUNCATION_VALUE = 1e-25  # .pyx:12-13
# Cython narrows the Python float to C ``float`` before calling ``log`` (generated C:
# ``__Pyx_PyFloat_AsFloat`` then ``log(double)``), so log(trunc) is log((double)(float)1e-25).
LOG_TRUNCATION = float(np.log(np.float64(np.float32(TRUNCATION_VALUE))))


def cash(n_on, mu_on, truncation_value=TRUNCATION_VALUE):
    """stats/fit_statistics.py:29-79."""
    n_on = np.asanyarray(n_on)
    mu_on = np.asanyarray(mu_on)
    mu_on = np.where(mu_on <= truncation_value, truncation_value, mu_on)
    with np.errstate(divide="ignore", invalid="ignore"):
        stat = 2 * (mu_on - n_on * np.log(mu_on))
    return stat


def cash_sum(counts, npred):
    """.pyx:55-85 restated: 2*sum(npr' - [n>0] n log npr')."""
    counts = np.asarray(counts, dtype=np.float64).ravel()
    npred = np.asarray(npred, dtype=np.float64).ravel()
    above = npred > TRUNCATION_VALUE
    npr = np.where(above, npred, TRUNCATION_VALUE)
    with np.errstate(divide="ignore", invalid="ignore"):
        lognpr = np.where(above, np.log(np.where(above, npred, 1.0)), LOG_TRUNCATION)
    pos = counts > 0
    terms = npr - np.where(pos, counts * lognpr, 0.0)
    return 2 * float(np.sum(terms))


def weighted_cash_sum(counts, npred, weight):
    """.pyx:17-51 restated."""
    counts = np.asarray(counts, dtype=np.float64).ravel()
    npred = np.asarray(npred, dtype=np.float64).ravel()
    weight = np.asarray(weight, dtype=np.float64).ravel()
    above = npred > TRUNCATION_VALUE
    npr = np.where(above, npred, TRUNCATION_VALUE)
    with np.errstate(divide="ignore", invalid="ignore"):
        lognpr = np.where(above, np.log(np.where(above, npred, 1.0)), LOG_TRUNCATION)
    w = weight > 0
    terms = np.where(w, weight * npr, 0.0) - np.where(w & (counts > 0), weight * counts * lognpr, 0.0)
    return 2 * float(np.sum(terms))


def reference_module():
    """The reference's compiled Cython module (``oracle/_ref``) or None."""
    from . import build_ref

    try:
        return build_ref.load()
    except Exception:  # toolchain missing on this box: restatement only
        return None


def stat_sum_dataset(counts, npred, mask=None, use_reference=True):
    """CashFitStatistic.stat_sum_dataset (fit_statistics.py:281-293): boolean gather then the compiled sum."""
    counts = np.asarray(counts)
    npred = np.asarray(npred)
    if mask is not None:
        mask = np.asarray(mask).astype("bool")
        counts, npred = counts[mask], npred[mask]
    counts = counts.astype(float)
    ref = reference_module() if use_reference else None
    if ref is not None:
        return float(ref.cash_sum_cython(np.ascontiguousarray(counts.ravel()),
                                         np.ascontiguousarray(npred.ravel(), dtype=float)))
    return cash_sum(counts.ravel(), npred.ravel())
