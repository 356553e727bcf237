"""Sustained (power-capped) rate of the scoring step at the headline shape: K steps back to back, clocks sampled.
usage: python tools/sustain.py [steps] ; knobs through the environment (B7_POST_DBG, B7_POST_PAIR, B7_KSTAR_OVERLAP ...)"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from bot7_b200 import _lib as L
from bot7_b200 import grids, models
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ctx = L.Context.default(0); lib = L.lib()
sob = grids.sobol({"size": bench.N_OBS + bench.M_STEP, "dims": bench.DIMS})
Xo = sob.generate({"size": bench.N_OBS, "dims": bench.DIMS})
y = bench.hartmann6(Xo); y = (y - y.mean()) / y.std()
f = models.GPFactors(Xo, y, bench.hyper_draws(bench.S_DRAWS, bench.DIMS))
grid = sob.generate_device(first=bench.N_OBS, count=bench.M_STEP)
def step():
    o_, b_, n_ = C.c_int64(), C.c_double(), C.c_int64()
    L.check(lib.b7_acq_score_range(f.handle, grid.handle, 0, bench.M_STEP, 0, 0.0, 0, -1.0, float(y.min()), None, C.byref(o_), C.byref(b_), C.byref(n_)))
    return o_.value, b_.value
for _ in range(3): r = step()
with bench.ClockSampler(0) as clk:
    ctx.timer_begin()
    for _ in range(steps): r = step()
    ms = ctx.timer_end() / steps
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("B7_")}, "ms_per_step": round(ms, 2), "cand_per_s": round(bench.M_STEP / ms * 1e3),
                  "tf_eq": round(bench.M_STEP * 32 * 4096.0 ** 2 / ms * 1e-9, 1), "clocks": clk.summary(), "argmax": r[0]}))
