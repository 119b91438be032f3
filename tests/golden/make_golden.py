"""Golden vectors of the reference's own data-free tests for this path, transcribed (file:line) into JSON.

The Python reference cannot be imported in this container (astropy / regions / iminuit are absent), so the
fixtures are the constants gammapy's test-suite asserts, not outputs of a reference run.  Re-run to regenerate
tests/golden/reference_kats.json; the paths are relative to /root/reference/gammapy.
"""
import json
import os

GOLDEN = {
    "cash": {
        "source": "stats/tests/test_fit_statistics.py:22-100 (Sherpa-derived per-bin values)",
        "mu_sig": [0.59752422, 9.13666449, 12.98288095, 5.56974565, 13.52509804, 11.81725635, 0.47963765,
                   11.17708176, 5.18504894, 8.30202394],
        "n_on": [0, 13, 7, 5, 11, 16, 0, 9, 3, 12],
        "cash": [1.19504844, -39.24635098872072, -9.925081055136996, -6.034002586236575, -30.249839537105466,
                 -55.39143500383233, 0.9592753, -21.095413867175516, 0.49542219758430406, -34.19193611846045],
    },
    "spectral": {
        "source": "modeling/models/tests/test_spectral.py:59-190 (TEST_MODELS)",
        "powerlaw": {"index": 2.3, "amplitude": 4.0, "reference": 1.0, "integral_1_10TeV": 2.9227116204223784},
        "ecpl": {"index": 1.6, "amplitude": 4.0, "reference": 1.0, "lambda_": 0.1, "val_at_2TeV": 1.080321705479446,
                 "integral_1_10TeV": 3.765838739678921},
        "logpar": {"alpha": 2.3, "amplitude": 4.0, "reference": 1.0, "beta": 0.5, "val_at_2TeV": 0.6387956571420305,
                   "integral_1_10TeV": 2.255689748270628},
    },
    "spatial": {
        "source": "modeling/models/tests/test_spatial.py:157-245, modeling/models/tests/test_cube.py:802-845",
        "disk_value": 261.263956,
        "integrate_nd_geom_pixel_12_10": [1.973745e-13, 9.161312e-14, 4.252304e-14],
        "evaluate_nd_geom_pixel_12_10": [2.278184e-07, 4.908198e-08, 1.057439e-08],
    },
    "dataset": {
        "source": "datasets/tests/test_map.py:1418-1445 (test_npred_psf_after_edisp)",
        "npred_sum": 129553.858658,
    },
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as fh:
        json.dump(GOLDEN, fh, indent=1)
    print(out)
