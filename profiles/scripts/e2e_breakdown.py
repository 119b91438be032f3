"""Where the end-to-end step (Datasets.stat_sum() with host Parameter objects) spends its time beyond the device step."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import ctypes as C
from gammapy_b200 import Datasets, configs, _lib
ds = configs.make_c3()
ds.fake(random_state=2)
dsets = Datasets([ds])
eng = dsets._get_engine()
pars = [p for p in eng.parameters if p.name in ("index", "alpha")]
v0 = [p.value for p in pars]
for k in range(20):
    dsets.stat_sum()
N = 400
def timeit(fn):
    t = time.perf_counter()
    for k in range(N):
        fn(k)
    return (time.perf_counter() - t) / N * 1e6
def full(k):
    for p, v in zip(pars, v0):
        p.value = v + (1e-3 if k % 2 == 0 else -1e-3)
    dsets.stat_sum()
def setters(k):
    for p, v in zip(pars, v0):
        p.value = v + (1e-3 if k % 2 == 0 else -1e-3)
vals = eng.values()
stat = np.empty(1)
def ccall(k):
    vals[0] += 1e-9 if k % 2 == 0 else -1e-9
    _lib.check(eng.ctx.lib.gpy_stat_sum(eng.ctx.handle, eng._handles, 1, vals.ctypes.data_as(C.c_void_p), 1, eng.n_par, stat.ctypes.data_as(C.c_void_p), None))
print("full step (setters + Datasets.stat_sum) %.1f us" % timeit(full))
print("50 Parameter setters            %.1f us" % timeit(setters))
print("Datasets._get_engine            %.1f us" % timeit(lambda k: dsets._get_engine()))
print("engine.values() (%d parameters)  %.1f us" % (eng.n_par, timeit(lambda k: eng.values())))
print("engine.stat_sum(values given)   %.1f us" % timeit(lambda k: eng.stat_sum(vals)))
print("bare gpy_stat_sum C call        %.1f us" % timeit(ccall))
