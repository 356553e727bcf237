"""Stage timings at a given shape (development aid)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b7_oracle as o
from bot7_b200 import _lib as L
lib = L.lib(); ctx = L.Context(0)
N, d, S, M = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (4096, 6, 32, 148 * 128 * 2))]
r = np.random.default_rng(1)
X = o.sobol_points(d, N + M); Xo, Xc = X[:N].copy(), X[N:].copy()
y = o.hartmann6(Xo) if d == 6 else o.ackley(Xo); y = (y - y.mean()) / y.std()
hyp = np.zeros((S, d + 3)); hyp[:, :d] = np.log(0.1) + r.random((S, d)) * (np.log(2) - np.log(0.1))
hyp[:, d] = 0.5 * (r.random(S) - 0.5); hyp[:, d + 1] = 0.5 * np.log(1e-2); hyp[:, d + 2] = 0.1 * (r.random(S) - 0.5)
ctx.set_profiling(True)
for rep in range(2):
    ctx.reset_timers()
    gp = C.c_void_p(); info = (C.c_int * S)(); logml = np.zeros(S); jit = np.zeros(S)
    t0 = time.time()
    L.check(lib.b7_gp_fit(ctx.handle, 0, L.dptr(Xo), L.dptr(y), N, d, L.dptr(hyp), S, d + 3, 0, 0, C.byref(gp), info, L.dptr(logml), L.dptr(jit)), "fit")
    t1 = time.time()
    grid = C.c_void_p(); L.check(lib.b7_grid_from_host(ctx.handle, L.dptr(Xc), M, d, C.byref(grid)))
    am = C.c_int64(); amo = C.c_int64(); best = C.c_double(); nn = C.c_int64()
    t2 = time.time()
    L.check(lib.b7_acq_score(gp, grid, 0, 0.0, 0, -1.0, float(y.min()), None, C.byref(am), C.byref(amo), C.byref(best), C.byref(nn)), "acq")
    t3 = time.time()
    st = ctx.stage_times()
    print(f"rep {rep}: fit wall {1e3*(t1-t0):.1f} ms, acq wall {1e3*(t3-t2):.1f} ms ({M/(t3-t2):.0f} cand/s), info {sum(info)} argmax {am.value}")
    print("  stages:", {k: (round(v[0], 3), v[1]) for k, v in st.items() if v[1]})
    Np = (N + 127) // 128 * 128
    print(f"  potrf {S*N**3/3/st['potrf'][0]*1e-9:.2f} TF  trtri {S*N**3/3/st['trtri'][0]*1e-9:.2f} TF  posterior {M*S*float(N)**2/st['posterior'][0]*1e-9:.2f} TF "
          f" kstar {M*S*Np*8/st['kstar'][0]*1e-6:.1f} GB/s  kbuild {S*Np*Np*8/st['kbuild'][0]*1e-6:.1f} GB/s")
    lib.b7_grid_free(grid); lib.b7_gp_free(gp)
