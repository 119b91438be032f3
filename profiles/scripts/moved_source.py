"""Per-call wall time of an evaluation with one source moved (spatial model + PSF stencil of that source + descriptor
rebuild + fold), the 50 sources of C3 in turn; prints the per-source-type medians."""
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
kinds = [type(m.spatial_model).__name__ for m in ds.models if hasattr(m, "spatial_model") and m.spatial_model is not None]
stat_dev = torch.zeros(1, 1, dtype=torch.float64, device=eng.ctx.torch_device)
def move(k, rep):
    v = base.copy()
    v[lon_cols[k]] += 1e-3 * (1 + rep)
    eng.stat_sum_device(v, stat_dev)
for k in range(len(lon_cols)):
    move(k, -2)
stream.synchronize()
res = {}
for rep in range(4):
    for k in range(len(lon_cols)):
        t0 = time.perf_counter()
        move(k, rep)
        stream.synchronize()
        res.setdefault(kinds[k], []).append((time.perf_counter() - t0) * 1e3)
for kind, v in res.items():
    v = np.array(v)
    print(f"{kind:24s} n={len(v):3d} median {np.median(v):.3f} ms  mean {v.mean():.3f}  max {v.max():.3f}")
allv = np.concatenate([np.array(v) for v in res.values()])
print(f"all: median {np.median(allv):.3f} mean {allv.mean():.3f} ms; steady-state step for comparison:")
t0 = time.perf_counter()
for k in range(50):
    eng.stat_sum_device(base, stat_dev)
stream.synchronize()
print(f"  {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
