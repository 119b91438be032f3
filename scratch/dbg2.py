import sys, numpy as np, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, "/root/repo")
from gammapy_b200 import Datasets, Map, configs
ds = configs.make_c1(width=(2.0, 2.0), mask_radius=0.9)
ds.fake(0)
dss = Datasets([ds])
print("B=1", dss.stat_sum(), flush=True)
v, pl = dss.parameter_vector()
cols = {p.name: i for i, p in enumerate(pl)}
for B in (2, 5, 7, 12):
    g = np.tile(v, (B, 1))
    g[:, cols["lon_0"]] += np.linspace(0, 0.02, B)
    print("B spatial", B, dss.stat_sum_batch(g)[:2], flush=True)
g = np.tile(v, (3, 1)); g[:, cols["sigma"]] += [0, 0.3, -0.25]
print("sigma", dss.stat_sum_batch(g), flush=True)
c2 = configs.make_c2(width=(2.0, 2.0), mask_radius=0.9); c2.fake(1)
d2 = Datasets([c2]); v2, _ = d2.parameter_vector()
print("c2 B=1", d2.stat_sum(), flush=True)
print("c2 B=40", d2.stat_sum_batch(np.tile(v2, (40, 1)))[:2], flush=True)
