"""oracle/wstat.py against the reference's own WStat goldens (stats/tests/test_fit_statistics.py)."""
import json
import os

import numpy as np
from numpy.testing import assert_allclose

from oracle import wstat as ow

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wstat_kats.json")))


def test_wstat_per_bin_goldens():
    got = ow.wstat(KATS["n_on"], KATS["n_off"], KATS["alpha"], KATS["mu_sig"], extra_terms=True)
    assert_allclose(got, KATS["wstat"])


def test_wstat_corner_cases():
    """test_fit_statistics.py:163-205."""
    assert_allclose(ow.wstat(0, 5, 0.5, 2.3), 2 * (2.3 + 5 * np.log(1.5)))
    assert_allclose(ow.get_wstat_mu_bkg(0, 5, 0.5, 2.3), 5 / 1.5)
    assert_allclose(ow.wstat(9, 0, 0.5, 2.3), -2 * (2.3 / 0.5 + 9 * np.log(0.5 / 1.5)))
    assert_allclose(ow.get_wstat_mu_bkg(9, 0, 0.5, 2.3), 9 / 1.5 - 2.3 / 0.5)
    assert_allclose(ow.wstat(5, 0, 0.5, 5.3), 2 * (5.3 + 5 * (np.log(5) - np.log(5.3) - 1)))
    assert_allclose(ow.get_wstat_mu_bkg(5, 0, 0.5, 5.3), 0)


def test_wstat_dataset_sums():
    d = KATS["dataset"]
    assert_allclose(ow.stat_sum_dataset(d["counts"], d["counts_off"], d["alpha"], d["npred_signal"]), d["stat_sum_nomask"])
    assert_allclose(ow.stat_sum_dataset(d["counts"], d["counts_off"], d["alpha"], d["npred_signal"], d["mask"]),
                    d["stat_sum_mask"])
    assert ow.stat_sum_dataset(d["counts"], d["counts_off"], d["alpha"], d["npred_signal"], [False] * 3) == 0
