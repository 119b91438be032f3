"""Wall vs device time of the batched C3 calls (scan of one parameter, full FD gradient)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from gammapy_b200 import Datasets

ds = bench.build_dataset("c3", 0)
datasets = Datasets([ds])
ds.fake(random_state=2)
engine = datasets._get_engine()
ctx = engine.ctx
stream = ctx.stream
base = engine.values()
cols = bench.spectral_columns(engine)

def timed(grid, n=3, label=""):
    sd = torch.zeros(grid.shape[0], 1, dtype=torch.float64, device=ctx.torch_device)
    engine.stat_sum_device(grid, sd)
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for k in range(n):
        engine.stat_sum_device(grid, sd)
    t_host = time.perf_counter() - t0
    e1.record(stream)
    stream.synchronize()
    t1 = time.perf_counter() - t0
    print(label, "B", grid.shape[0], "device ms", e0.elapsed_time(e1) / n, "host-issue ms", t_host * 1e3 / n, "wall ms", t1 * 1e3 / n, flush=True)

B = 32
grid = np.tile(base, (B, 1)); grid[:, cols[0]] += np.linspace(-0.3, 0.3, B)
timed(grid, label="scan")
for sel, label in ((list(cols), "grad-spectral"), ([i for i, p in enumerate(engine.parameters) if p.name in ("lon_0", "lat_0")], "grad-position")):
    grid = np.tile(base, (1 + 2 * len(sel), 1))
    for j, c in enumerate(sel):
        h = 1e-3
        grid[1 + 2 * j, c] += h; grid[2 + 2 * j, c] -= h
    timed(grid, label=label)
