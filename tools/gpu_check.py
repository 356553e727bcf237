"""Quick GPU-vs-oracle diagnostics through the raw C ABI (development aid; the real parity tests
are tests/test_gpu_*.py)."""
import ctypes as C
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b7_oracle as o  # noqa: E402
from bot7_b200 import _lib as L  # noqa: E402

lib = L.lib()
ctx = L.Context(0)
rng = np.random.default_rng(0)


def rel(a, b, floor=0.0):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor + 1e-300)))


# ---- sobol
for dims, first, count in [(6, 1, 65536), (2, 1, 20000), (20, 5000, 100000), (39, 4090, 9000)]:
    out = np.empty((count, dims))
    L.check(lib.b7_sobol_generate(ctx.handle, dims, first, count, None, None, L.dptr(out), None), "sobol")
    ref = o.sobol_numerators(dims, first, count).astype(np.float64) * 2.0 ** -30
    print("sobol", dims, first, count, "equal:", np.array_equal(out, ref))
num = (out * 2 ** 30).astype("<u4")
mins = np.linspace(-1, 0.5, 6); maxes = np.linspace(1, 3.3, 6)
out = np.empty((1000, 6))
L.check(lib.b7_sobol_generate(ctx.handle, 6, 1, 1000, L.dptr(mins), L.dptr(maxes), L.dptr(out), None))
print("sobol affine equal:", np.array_equal(out, o.sobol_points(6, 1000, 1, mins, maxes)))

# ---- score moments
S, M = 5, 100003
mean = rng.normal(size=(S, M)); var = rng.random((S, M)) ** 2
var[0, :10] = 0.0; mean[1, 5] = np.nan; var[2, 7] = -1.0
fmin = -0.3
for kind, name in [(0, "EI"), (1, "CB")]:
    sc = np.empty(M); am = C.c_int64(); best = C.c_double(); nn = C.c_int64()
    trade = 0.0 if kind == 0 else 1.0
    L.check(lib.b7_score_moments(ctx.handle, kind, L.dptr(mean), L.dptr(var), S, M, trade, 0, -1.0, fmin, L.dptr(sc),
                                 C.byref(am), C.byref(best), C.byref(nn)))
    per = [o.ei_compute(mean[s], var[s], fmin, trade) if kind == 0 else o.cb_compute(mean[s], var[s], trade) for s in range(S)]
    ref = o.mc_average(per)
    b, i, n = o.argmax_first(ref)
    ok = np.isfinite(ref)
    print(name, "bit-equal frac:", float(np.mean(sc[ok] == ref[ok])), "max rel:", rel(sc[ok], ref[ok]), "nan same:",
          np.array_equal(np.isnan(sc), np.isnan(ref)), "argmax", am.value, i, "best", best.value, b, "nan", nn.value, n)

# ---- GP fit / predict / acquisition
def make_problem(N, d, S, M, noise, seed=1):
    r = np.random.default_rng(seed)
    X = o.sobol_points(d, N + M)
    perm = r.permutation(N + M)
    Xo, Xc = X[np.sort(perm[:N])], X[np.sort(perm[N:])]
    y = o.hartmann6(Xo) if d == 6 else np.sin(3 * Xo.sum(1))
    y = (y - y.mean()) / y.std()
    hyp = np.zeros((S, d + 3))
    hyp[:, :d] = np.log(0.1) + r.random((S, d)) * (np.log(2) - np.log(0.1))
    hyp[:, d] = 0.5 * (r.random(S) - 0.5)
    hyp[:, d + 1] = 0.5 * np.log(noise)
    hyp[:, d + 2] = 0.1 * (r.random(S) - 0.5)
    return Xo, y, hyp, Xc


for (N, d, S, M, noise, kern) in [(50, 2, 3, 2000, 1e-2, 0), (300, 6, 3, 5000, 1e-2, 0), (300, 6, 2, 3000, 1e-2, 1),
                                  (1000, 6, 2, 20000, 1e-2, 0), (515, 20, 2, 4000, 1e-6, 0)]:
    Xo, y, hyp, Xc = make_problem(N, d, S, M, noise)
    gp = C.c_void_p(); info = (C.c_int * S)(); logml = np.zeros(S); jit = np.zeros(S)
    t0 = time.time()
    L.check(lib.b7_gp_fit(ctx.handle, kern, L.dptr(Xo), L.dptr(y), N, d, L.dptr(hyp), S, d + 3, 0, 0, C.byref(gp), info,
                          L.dptr(logml), L.dptr(jit)), "gp_fit")
    t1 = time.time()
    fits = [o.gp_fit(Xo, y, hyp[s], kern) for s in range(S)]
    print(f"GP N={N} d={d} S={S} kern={kern} noise={noise:g}: info {list(info)} jitter {jit.tolist()} fit {1e3*(t1-t0):.1f} ms")
    print("   logml rel:", rel(logml, np.array([f['logml'] for f in fits])))
    # inverse factor check
    Li = np.empty((N, N)); L.check(lib.b7_gp_read_factor(gp, 0, L.dptr(Li)))
    Linv_ref = np.linalg.inv(fits[0]["L"])
    print("   Linv max abs err (scaled):", float(np.max(np.abs(np.tril(Li) - Linv_ref)) / np.max(np.abs(Linv_ref))))
    for s in range(S):
        mu = np.empty(M); va = np.empty(M)
        L.check(lib.b7_gp_predict(gp, s, L.dptr(Xc), M, L.dptr(mu), L.dptr(va)), "predict")
        mr, vr = o.gp_predict(fits[s], Xc)
        print(f"   draw {s}: mean rel(|mu|>=1e-3) {rel(mu, mr, 1e-3):.2e} abs {np.max(np.abs(mu-mr)):.2e}  var rel {rel(va, vr, 1e-12):.2e} "
              f"abs/sf2 {np.max(np.abs(va-vr))/fits[s]['sf2']:.2e} minvar {vr.min():.2e}")
    grid = C.c_void_p(); L.check(lib.b7_grid_from_host(ctx.handle, L.dptr(Xc), M, d, C.byref(grid)))
    for kind in (0, 1):
        sc = np.empty(M); am = C.c_int64(); amo = C.c_int64(); best = C.c_double(); nn = C.c_int64()
        trade = 0.0 if kind == 0 else 1.0
        L.check(lib.b7_acq_score(gp, grid, kind, trade, 0, -1.0, float(y.min()), L.dptr(sc), C.byref(am), C.byref(amo),
                                 C.byref(best), C.byref(nn)), "acq")
        ref = o.acquisition(Xo, y, hyp, Xc, kern, False, kind)
        print(f"   acq kind {kind}: score rel {rel(sc, ref['score'], 1e-12):.2e} abs {np.max(np.abs(sc-ref['score'])):.2e} argmax {am.value} vs {ref['idx']} "
              f"best {best.value:.12g} vs {ref['best']:.12g}")
    lib.b7_grid_free(grid); lib.b7_gp_free(gp)

# jitter path: duplicate points, tiny noise -> not PD
N, d, S = 200, 3, 2
Xo = rng.random((N, d)); Xo[100:] = Xo[:100]
y = rng.normal(size=N)
hyp = np.zeros((S, d + 3)); hyp[:, d + 1] = 0.5 * np.log(1e-18); hyp[:, :d] = np.log(0.5)
gp = C.c_void_p(); info = (C.c_int * S)(); logml = np.zeros(S); jit = np.zeros(S)
L.check(lib.b7_gp_fit(ctx.handle, 0, L.dptr(Xo), L.dptr(y), N, d, L.dptr(hyp), S, d + 3, 0, 0, C.byref(gp), info, L.dptr(logml), L.dptr(jit)))
fr = o.gp_fit(Xo, y, hyp[0], 0)
print("jitter path: gpu", jit.tolist(), list(info), "oracle", fr["jitter"], fr["iters"], "logml", logml[0], fr["logml"])
lib.b7_gp_free(gp)
print("launches", ctx.launch_count())
