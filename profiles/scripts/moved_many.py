"""Evaluations that move many sources at once, with the spatial chains of all moved sources in three launches (default)
and with three launches per source (GPY_SPATIAL_BATCH=0, read at context creation): every source of C3 shifted (50 spatial chains), one source shifted, and a central-difference gradient
over the 100 position parameters at a new point (201 vectors, 200 displaced cubes to build).  Run once per setting; prints wall time per call after a synchronize."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from gammapy_b200 import Datasets, configs

ds = configs.make_c3()
ds.fake(random_state=2)
dsets = Datasets([ds])
eng = dsets._get_engine()
base = eng.values()
stream = eng.ctx.stream
lon_cols = [i for i, p in enumerate(eng.parameters) if p.name == "lon_0"]
stat_dev = torch.zeros(1, 1, dtype=torch.float64, device=eng.ctx.torch_device)


def timed(fn, reps):
    out = []
    for r in range(reps):
        t0 = time.perf_counter()
        fn(r)
        stream.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
    return np.array(out)


def move_all(r):
    v = base.copy()
    v[lon_cols] += 1e-3 * (1 + r)
    eng.stat_sum_device(v, stat_dev)


def move_one(r):
    v = base.copy()
    v[lon_cols[r % len(lon_cols)]] += 1e-3 * (2 + r)
    eng.stat_sum_device(v, stat_dev)


free = [i for i, p in enumerate(eng.parameters) if p.name in ("lon_0", "lat_0")]
batch_out = torch.zeros(2 * len(free) + 1, 1, dtype=torch.float64, device=eng.ctx.torch_device)


def fd_gradient(r):
    x = base.copy()
    x[lon_cols] += 1e-4 * r
    V = np.tile(x, (2 * len(free) + 1, 1))
    for k, i in enumerate(free):
        h = 1e-3
        V[1 + 2 * k, i] += h
        V[2 + 2 * k, i] -= h
    eng.stat_sum_device(V, batch_out)


label = "batched spatial launches " + os.environ.get("GPY_SPATIAL_BATCH", "1")
for name, fn, reps in (("all 50 sources moved", move_all, 12), ("one source moved", move_one, 50),
                       (f"position gradient ({2 * len(free) + 1} vectors)", fd_gradient, 6)):
    fn(-1)
    stream.synchronize()
    t = timed(fn, reps)
    print(f"{label}: {name:32s} median {np.median(t):.3f} ms  min {t.min():.3f}  max {t.max():.3f}")
move_all(100)
stream.synchronize()
ref = float(stat_dev.cpu()[0, 0])
print(f"{label}: statistic after the last move {ref!r}")
