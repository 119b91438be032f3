"""Golden TS map of a small synthetic field: inputs + the outputs of the reference's compiled statistics
(``fit_statistics_cython.pyx`` compiled unmodified into ``oracle/_ref``) driven by ``scipy.optimize.brentq`` in the
call order of ``estimators/map/ts.py`` (``oracle.tsmap``).  Needs /root/reference (to build oracle/_ref) and scipy.

    python tests/golden/make_ts_map_golden.py   ->   tests/golden/ts_map_small.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def make_inputs(seed=7, n_planes=2, ny=22, nx=26, k=7):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[:ny, :nx]
    g = np.exp(-0.5 * ((np.arange(k) - k // 2) / 1.2) ** 2)
    kernel = np.stack([np.outer(g, g) / np.outer(g, g).sum(),
                       np.outer(g ** 2, g ** 2) / np.outer(g ** 2, g ** 2).sum()])[:n_planes]
    exposure = 1e3 * (1 + 0.3 * np.cos(xx / 9.0))[None] * np.array([1.0, 0.6])[:n_planes, None, None]
    background = (2.0 + 0.05 * yy)[None] * np.array([1.0, 0.3])[:n_planes, None, None]
    source = np.zeros((ny, nx))
    source[9, 12] = 0.08
    source[15, 5] = 0.02
    npred = background.copy()
    for e in range(n_planes):
        pad = np.pad(source * exposure[e], k // 2)
        for dy in range(k):
            for dx in range(k):
                npred[e] += kernel[e, dy, dx] * pad[k - 1 - dy:k - 1 - dy + ny, k - 1 - dx:k - 1 - dx + nx]
    counts = rng.poisson(npred).astype(float)
    # a strip with no exposure, no background, no counts (padded region: bins dropped as invalid)
    for a in (counts, background, exposure):
        a[:, :, :2] = 0.0
    # a corner where nothing was counted although background is expected (counts.sum() <= 0 branch)
    counts[:, 17:, 19:] = 0.0
    mask = np.ones((ny, nx), dtype=bool)
    mask[:, :2] = False
    mask[3:6, 20:23] = False
    return dict(counts=counts, background=background, exposure=exposure, kernel=kernel, mask=mask)


def main():
    from oracle import cash, tsmap

    if cash.reference_module() is None:
        raise SystemExit("oracle/_ref is not built: run oracle/build_ref.py where /root/reference is available")
    inputs = make_inputs()
    out = tsmap.ts_map(rtol=0.01, max_niter=100, n_sigma=1.0, use_reference=True, use_scipy=True, **inputs)
    tight = tsmap.ts_map(rtol=1e-10, max_niter=100, n_sigma=1.0, use_reference=True, use_scipy=True, **inputs)
    np.savez_compressed(os.path.join(HERE, "ts_map_small.npz"), **inputs, **{f"out_{k}": v for k, v in out.items()},
                        **{f"tight_{k}": v for k, v in tight.items()})
    print({k: float(np.nanmax(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
