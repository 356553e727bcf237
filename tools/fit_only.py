"""Batched fit only (for ncu launch lists): N, d, S from argv."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b7_oracle as o
from bot7_b200 import _lib as L, models
N, d, S = [int(x) for x in sys.argv[1:4]]
X = o.sobol_points(d, N); y = o.hartmann6(X) if d == 6 else o.ackley(X); y = (y - y.mean()) / y.std()
r = np.random.default_rng(1)
hyp = np.zeros((S, d + 3)); hyp[:, :d] = np.log(0.1) + r.random((S, d)) * (np.log(2) - np.log(0.1)); hyp[:, d + 1] = 0.5 * np.log(1e-2)
ctx = L.Context.default(0); ctx.set_profiling(True)
f = models.GPFactors(X, y, hyp)
for i in range(2):
    ctx.reset_timers(); f.refit(hyp + 0.01 * i)
    print({k: v for k, v in ctx.stage_times().items() if v[1]})
