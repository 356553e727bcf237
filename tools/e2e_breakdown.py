"""Host-side breakdown of one end-to-end step (fit from host arrays, grid upload, acquisition with read-back, frees)."""
import ctypes as C, os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b7_oracle as o
from bot7_b200 import _lib as L, models, grids
lib = L.lib(); ctx = L.Context.default(0)
N, d, S, M = 4096, 6, 32, 148 * 128 * 2
r = np.random.default_rng(1)
X = o.sobol_points(d, N + M); Xo, Xc = X[:N].copy(), X[N:].copy()
y = o.hartmann6(Xo); y = (y - y.mean()) / y.std()
hyp = np.zeros((S, d + 3)); hyp[:, :d] = np.log(0.1) + r.random((S, d)) * (np.log(2) - np.log(0.1))
hyp[:, d] = 0.5 * (r.random(S) - 0.5); hyp[:, d + 1] = 0.5 * np.log(1e-2); hyp[:, d + 2] = 0.1 * (r.random(S) - 0.5)
rows = []
for it in range(6):
    t = [time.perf_counter()]
    f = models.GPFactors(Xo, y, hyp); t.append(time.perf_counter())
    g = grids.DeviceGrid.from_host(Xc, ctx); t.append(time.perf_counter())
    score = np.empty(M); am, amo, b_, n_ = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
    L.check(lib.b7_acq_score(f.handle, g.handle, L.SCORE_EI, 0.0, 0, -1.0, float(y.min()), L.dptr(score), C.byref(am), C.byref(amo), C.byref(b_), C.byref(n_)))
    t.append(time.perf_counter())
    f.free(); g.free(); t.append(time.perf_counter())
    rows.append([round(1e3 * (t[i + 1] - t[i]), 2) for i in range(4)])
print(json.dumps({"columns": ["fit_ms", "grid_upload_ms", "acq_ms", "free_ms"], "iterations": rows}))
