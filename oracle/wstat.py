"""Oracle: WStat (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows (paths relative to /root/reference/gammapy):

* ``wstat``                 stats/fit_statistics.py:142-218
* ``get_wstat_mu_bkg``      stats/fit_statistics.py:221-236
* ``get_wstat_gof_terms``   stats/fit_statistics.py:239-252
* ``WStatFitStatistic``     stats/fit_statistics.py:332-361 (np.nan_to_num per bin, mask gather, sum)

Pinned by the reference's own goldens (tests/golden/wstat_kats.json, copied from stats/tests/test_fit_statistics.py:
20-116 Sherpa-derived per-bin values, :163-205 corner cases, :331-371 dataset sums)."""
import numpy as np


def get_wstat_mu_bkg(n_on, n_off, alpha, mu_sig):
    n_on, n_off, alpha, mu_sig = (np.asanyarray(a, dtype=np.float64) for a in (n_on, n_off, alpha, mu_sig))
    c = alpha * (n_on + n_off) - (1 + alpha) * mu_sig
    d = np.sqrt(c ** 2 + 4 * alpha * (alpha + 1) * n_off * mu_sig)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (c + d) / (2 * alpha * (alpha + 1))


def get_wstat_gof_terms(n_on, n_off):
    term = np.zeros(np.shape(n_on))
    with np.errstate(divide="ignore", invalid="ignore"):
        term1 = -n_on * (1 - np.log(n_on))
        term2 = -n_off * (1 - np.log(n_off))
    term = term + np.where(n_on == 0, 0, term1)
    term = term + np.where(n_off == 0, 0, term2)
    return 2 * term


def wstat(n_on, n_off, alpha, mu_sig, mu_bkg=None, extra_terms=True):
    n_on, n_off, alpha, mu_sig = (np.asanyarray(a, dtype=np.float64) for a in (n_on, n_off, alpha, mu_sig))
    if mu_bkg is None:
        mu_bkg = get_wstat_mu_bkg(n_on, n_off, alpha, mu_sig)
    term1 = mu_sig + (1 + alpha) * mu_bkg
    with np.errstate(divide="ignore", invalid="ignore"):
        term2_ = -n_on * np.log(mu_sig + alpha * mu_bkg)
    term2 = np.where(n_on == 0, 0, term2_)
    with np.errstate(divide="ignore", invalid="ignore"):
        term3_ = -n_off * np.log(mu_bkg)
    term3 = np.where(n_off == 0, 0, term3_)
    stat = 2 * (term1 + term2 + term3)
    if extra_terms:
        stat = stat + get_wstat_gof_terms(n_on, n_off)
    return stat


def stat_array_dataset(counts, counts_off, alpha, npred_signal):
    return np.nan_to_num(wstat(counts, counts_off, alpha, npred_signal))


def stat_sum_dataset(counts, counts_off, alpha, npred_signal, mask=None):
    arr = stat_array_dataset(counts, counts_off, alpha, npred_signal)
    if mask is not None:
        arr = arr[np.asarray(mask, dtype=bool)]
    return float(np.sum(arr))
