"""Runs the TS-map kernel once on the bench field (for ncu): python profiles/ts_map_run.py [ny nx]"""
import sys
import time

sys.path.insert(0, ".")
import bench
from gammapy_b200.estimators import compute_ts_map

ny, nx = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (500, 2000)
counts, background, exposure, kernel = bench.ts_map_inputs(ny=ny, nx=nx)
compute_ts_map(counts[:, :64, :64], background[:, :64, :64], exposure[:, :64, :64], kernel)
for rep in range(3):
    t0 = time.perf_counter()
    res = compute_ts_map(counts, background, exposure, kernel, rtol=0.01, max_niter=100)
    print("seconds incl. copies", time.perf_counter() - t0, "mean niter", float(res["niter"].mean()), flush=True)
