"""Large-M sanity: 2^24 Sobol candidates (d = 20) generated on the device, scored in one call and in 3 uneven shards."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b7_oracle as o
from bot7_b200 import _lib as L, grids, models, parallel
N, d, S, M = 1024, 20, 2, 1 << 24
sob = grids.sobol({"size": N + M, "dims": d})
Xo = sob.generate({"size": N, "dims": d}); y = o.ackley(Xo); y = (y - y.mean()) / y.std()
r = np.random.default_rng(0)
hyp = np.zeros((S, d + 3)); hyp[:, :d] = np.log(0.5) + r.random((S, d)); hyp[:, d + 1] = 0.5 * np.log(1e-2)
f = models.GPFactors(Xo, y, hyp)
grid = sob.generate_device(first=N, count=M)
grid.remove(5); grid.remove(M - 10)
lib = L.lib()
t0 = time.perf_counter()
am, amo, best, nn = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
L.check(lib.b7_acq_score(f.handle, grid.handle, L.SCORE_EI, 0.0, 0, -1.0, float(y.min()), None, C.byref(am), C.byref(amo), C.byref(best), C.byref(nn)))
t1 = time.perf_counter()
print(f"M={M}: {t1 - t0:.2f} s, {M / (t1 - t0):.3e} cand/s; argmax {am.value} orig {amo.value} best {best.value} nan {nn.value}")
trips = []
for (r0, cnt) in ((0, 5_000_001), (5_000_001, 7_777_777), (12_777_778, M - 12_777_778)):
    o_, b_, n_ = C.c_int64(), C.c_double(), C.c_int64()
    L.check(lib.b7_acq_score_range(f.handle, grid.handle, r0, cnt, L.SCORE_EI, 0.0, 0, -1.0, float(y.min()), None, C.byref(o_), C.byref(b_), C.byref(n_)))
    trips.append((b_.value, o_.value, n_.value))
print("sharded:", parallel.combine_argmax(trips), "equal:", parallel.combine_argmax(trips) == (best.value, amo.value, nn.value))
# spot check the winner against the oracle
x = grid.read(amo.value - 1, 1)
ref = o.acquisition(Xo, y, hyp, x, 0, False, o.SCORE_EI)
print("winner score oracle", ref["score"][0], "rel err", abs(ref["score"][0] - best.value) / best.value)
assert np.array_equal(x, o.sobol_points(d, 1, skip=1 + N + amo.value - 1))
print("ok")
