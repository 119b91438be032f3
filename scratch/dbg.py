import sys, numpy as np
sys.path.insert(0, "/root/repo")
from gammapy_b200 import Datasets, Map, configs
ds = configs.make_c1(width=(2.0, 2.0), mask_radius=0.9)
ds.fake(0)
print("fake ok", flush=True)
dss = Datasets([ds])
print("B=1", dss.stat_sum(), flush=True)
v, _ = dss.parameter_vector()
for B in (2, 3, 8, 64):
    print("B", B, dss.stat_sum_batch(np.tile(v, (B, 1)))[:3], flush=True)
