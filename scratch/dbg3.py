import sys, faulthandler
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, "/root/repo")
import logging
from gammapy_b200 import _lib
ctx = _lib.Context.get(0)
import tests.test_gpu_fit as t
print("start", flush=True)
t.test_stat_profile_and_surface_batched(ctx)
print("profile ok", flush=True)
t.test_fit_gpu_vs_oracle_same_optimiser(ctx, "exact64", 1e-3)
print("fit ok", flush=True)
