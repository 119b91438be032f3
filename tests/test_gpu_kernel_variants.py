"""The fold + Cash path has three staged kernels: the 64-pixel sub-tile kernel (default when eligible), the 128-pixel
kernel (GPY_FOLD_T64=0; also what launches with two edisp groups or long axes use) and the double-buffered 64-pixel
variant (GPY_FOLD_T64=2), plus the any-shape kernel (GPY_FOLD_KERNEL=2).  The choice is made per context from the
environment, so every variant runs in its own process on the same seeded dataset; all must agree with the oracle and
with each other (the per-tile partial sums are added in the same fixed order: only the in-tile order differs)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
sys.path.insert(0, %r)
import numpy as np
from gammapy_b200 import Map, configs
from oracle import adapter
out = {}
for key, make in (("c2", lambda: configs.make_c2(width=(4.0, 4.0), mask_radius=1.2)),
                  ("c3s", lambda: configs.make_c3(width=(6.0, 3.0), n_sources=9))):
    ds = make()
    de = adapter.evaluator_for(ds, conv="exact64")
    ds.counts = Map.from_geom(ds._geom, data=de.fake(4), dtype=float)
    stat = ds.stat_sum()
    npred = ds.npred().data
    want = de.npred()
    eng = ds._get_engine()
    grid = np.tile(eng.values(), (3, 1))
    col = [i for i, p in enumerate(eng.parameters) if p.name in ("index", "alpha")][0]
    grid[:, col] += [0.0, 0.05, 0.1]
    scan = [float(v) for v in eng.stat_sum(grid)]
    ds.storage = "compact"  # (rebuilds the HBM mirror: the engine above is stale from here on)
    out[key] = {"stat": stat, "oracle": de.stat_sum(), "rel": float(np.max(np.abs(npred - want) / np.maximum(want, 1e-300))),
                "scan": scan, "compact": ds.stat_sum()}
print("RESULT " + json.dumps(out))
""" % ROOT


def _run(env):
    e = dict(os.environ)
    e.update(env)
    res = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=e, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_every_staged_kernel_variant_agrees(ctx):
    runs = {name: _run(env) for name, env in (("t64", {"GPY_FOLD_T64": "1"}), ("t128", {"GPY_FOLD_T64": "0"}),
                                              ("t64-double-buffered", {"GPY_FOLD_T64": "2"}),
                                              ("any-shape", {"GPY_FOLD_KERNEL": "2"}))}
    for name, out in runs.items():
        for key, r in out.items():
            assert abs(r["stat"] - r["oracle"]) < 1e-3, (name, key, r)
            assert r["rel"] < 1e-5, (name, key, r)
            assert r["compact"] == r["stat"], (name, key, r)       # same counts and mask, other storage format
            assert r["scan"][0] == r["stat"], (name, key, r)       # batched call, vector 0 = the single evaluation
    ref = runs["t64"]
    for name, out in runs.items():
        for key, r in out.items():
            assert abs(r["stat"] - ref[key]["stat"]) < 1e-6 * abs(r["stat"]) + 1e-6, (name, key)


@pytest.mark.parametrize("n_reco,n_true", [(36, 70), (12, 66), (34, 40)])
def test_long_axes_use_the_128_pixel_kernel(ctx, n_reco, n_true):
    """More than 32 reco bins (two sweeps over the reco chunks) or more than 64 true bins (three phase-1 passes): outside
    the 64-pixel kernel's shapes, served by the 128-pixel kernel; per-bin npred and the statistic against the oracle."""
    import numpy as np

    from gammapy_b200 import Map, configs
    from oracle import adapter

    ds = configs.make_dataset(width=(3.0, 2.0), n_reco=n_reco, n_true=n_true, e_reco=(0.1, 100.0), e_true=(0.03, 300.0),
                              mask_radius=0.8, name="long")
    ds.models = configs.c2_models("long")
    de = adapter.evaluator_for(ds, conv="exact64")
    got, want = ds.npred().data, de.npred()
    assert np.max(np.abs(got - want) / np.maximum(want, 1e-300)) < 1e-5
    ds.counts = Map.from_geom(ds._geom, data=de.fake(9), dtype=float)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3


@pytest.mark.parametrize("proj,skydir", [("TAN", (30.0, 40.0)), ("TAN", (300.0, -62.0)), ("CAR", (83.63, 22.01)),
                                         ("CAR", (266.4, -29.0))])
def test_off_equator_dataset_vs_oracle(ctx, proj, skydir):
    """A dataset on a TAN (gnomonic) or CAR (oblique plate carree, what WcsGeom.create(skydir=...) builds off the
    equator) geometry: model coordinates, solid angles, cutouts and the PSF kernel geometry follow the projection on the
    device (pix_to_world_deg), on the host (world_to_pix) and in the oracle (oracle/wcs.py, written from the FITS WCS
    paper in its other form); five sources of every spatial type."""
    import numpy as np

    from gammapy_b200 import Map, configs
    from gammapy_b200.modeling.models import FoVBackgroundModel
    from oracle import adapter

    ds = configs.make_dataset(width=(4.0, 3.0), skydir=skydir, proj=proj, mask_radius=1.3, name="tan")
    models = configs.c2_models("tan")
    for m in models:
        if not isinstance(m, FoVBackgroundModel):  # the ring of sources, re-centred on the field
            # (+ a few thousandths of a degree: a source exactly on the reference meridian or a whole number of pixels
            # from the reference parallel of an even-sized image sits on a pixel edge, where the cutout's rounding is
            # decided by the last bit of the rotation -- in wcslib as much as here)
            m.spatial_model.lon_0.value = skydir[0] + (m.spatial_model.lon_0.value + 0.0043) / np.cos(np.radians(skydir[1]))
            m.spatial_model.lat_0.value = skydir[1] + m.spatial_model.lat_0.value + 0.0031
    ds.models = models
    de = adapter.evaluator_for(ds, conv="exact64")
    got, want = ds.npred().data, de.npred()
    assert np.max(np.abs(got - want) / np.maximum(want, 1e-300)) < 1e-5
    for m in models:
        if not isinstance(m, FoVBackgroundModel):
            st = ds.evaluator_state(m.name)
            ys, xs = de.evaluators[m.name].slices
            assert (st["x0"], st["y0"], st["cx"], st["cy"]) == (xs.start, ys.start, xs.stop - xs.start, ys.stop - ys.start)
    ds.counts = Map.from_geom(ds._geom, data=de.fake(6), dtype=float)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3
    # a source moves: new cutout, same agreement
    models[1].spatial_model.lon_0.value += 0.4
    models[1].spatial_model.lat_0.value -= 0.3
    adapter.refresh_models(de, ds)
    assert abs(ds.stat_sum() - de.stat_sum()) < 1e-3


SIDE_SCRIPT = r"""
import json, sys
sys.path.insert(0, %r)
import numpy as np
from gammapy_b200 import Map, configs
from oracle import adapter
ds = configs.make_c3(width=(6.0, 3.0), n_sources=9)
de = adapter.evaluator_for(ds, conv="exact64")
ds.counts = Map.from_geom(ds._geom, data=de.fake(4), dtype=float)
first = ds.stat_sum()
for m in ds.models:                      # every source moves: nine spatial chains in one evaluation
    sm = getattr(m, "spatial_model", None)
    if sm is not None:
        sm.lon_0.value += 0.013
        sm.lat_0.value -= 0.007
moved = ds.stat_sum()
de2 = adapter.evaluator_for(ds, conv="exact64")
eng = ds._get_engine()
v = eng.values()
cols = [i for i, p in enumerate(eng.parameters) if p.name == "lon_0"]
grid = np.tile(v, (1 + len(cols), 1))
for k, c in enumerate(cols):
    grid[1 + k, c] += 2e-3               # batched call: nine displaced cubes built by the same three launches
scan = [float(x) for x in eng.stat_sum(grid)]
print("RESULT " + json.dumps({"first": first, "moved": moved, "oracle_moved": de2.stat_sum(), "scan": scan}))
""" % ROOT


def test_batched_spatial_launches_do_not_change_the_result(ctx):
    """The spatial chains (integrate_geom + PSF stencil) of all the sources an evaluation moved run as three launches
    (spatial_fill_batch_kernel, spatial_oversample_batch_kernel, psf_conv_batch_kernel: block -> job table): same bits as
    three launches per source (GPY_SPATIAL_BATCH=0), and the moved evaluation agrees with the oracle at the new
    positions."""
    runs = {}
    for name, env in (("side", {"GPY_SPATIAL_BATCH": "1"}), ("single", {"GPY_SPATIAL_BATCH": "0"})):
        e = dict(os.environ)
        e.update(env)
        res = subprocess.run([sys.executable, "-c", SIDE_SCRIPT], capture_output=True, text=True, env=e, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        runs[name] = json.loads([l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert runs["side"] == runs["single"]
    r = runs["side"]
    assert r["moved"] != r["first"]
    assert abs(r["moved"] - r["oracle_moved"]) < 1e-3 * max(1.0, abs(r["oracle_moved"]) * 1e-6), r
    assert r["scan"][0] == r["moved"]
    assert len(set(r["scan"])) == len(r["scan"])
