"""HBM-bound kernels on their own: Sobol generation, fused scoring pass, BLR scoring (config 4 shape)."""
import ctypes as C, os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bot7_b200 import _lib as L
lib = L.lib(); ctx = L.Context.default(0); ctx.set_profiling(True)
out = {}
# Sobol: config 5 shape slice (d=20) and d=6, device resident
for d, M in ((6, 1 << 24), (20, 1 << 24)):
    best = 1e9
    for rep in range(5):
        ctx.reset_timers(); g = C.c_void_p()
        L.check(lib.b7_sobol_generate(ctx.handle, d, 1, M, None, None, None, C.byref(g)))
        best = min(best, ctx.stage_times()["sobol"][0]); lib.b7_grid_free(g)
    out[f"sobol_d{d}_M{M}"] = {"ms": best, "GBps": M * d * 8 / best * 1e-6, "points_per_s": M / best * 1e3}
# scoring pass: S=32, M=4M moments resident is not exposed; time through b7_score_moments stage timer (kernel only)
S, M = 32, 1 << 21
r = np.random.default_rng(0); mean = r.normal(size=(S, M)); var = r.random((S, M))
for kind, nm in ((0, "ei"), (1, "cb")):
    best = 1e9
    for rep in range(3):
        ctx.reset_timers(); am = C.c_int64()
        L.check(lib.b7_score_moments(ctx.handle, kind, L.dptr(mean), L.dptr(var), S, M, 0.0 if kind == 0 else 1.0, 0, -1.0, -0.1, None, C.byref(am), None, None))
        best = min(best, ctx.stage_times()["score"][0])
    out[f"score_{nm}_S{S}_M{M}"] = {"ms": best, "GBps": M * (16 * S + 8) / best * 1e-6, "cand_per_s": M / best * 1e3}
# BLR config 4: N=20000, D=50, M=4M candidates (features resident)
N, D, Mb = 20000, 50, 1 << 22
Z0 = np.maximum(r.normal(size=(N, D)), 0); y = r.normal(size=N); Z1 = np.maximum(r.normal(size=(Mb, D)), 0)
hyp = np.array([[0.0, np.log(1e2), 0.0]])
h = C.c_void_p(); info = (C.c_int * 1)()
ctx.reset_timers()
L.check(lib.b7_blr_fit(ctx.handle, L.dptr(Z0), L.dptr(y), N, D, L.dptr(hyp), 1, C.byref(h), info))
out["blr_fit_N20000_D50"] = {"ms": ctx.stage_times()["blr"][0]}
g = C.c_void_p(); L.check(lib.b7_grid_from_host(ctx.handle, L.dptr(Z1), Mb, D, C.byref(g)))
best = 1e9
for rep in range(3):
    ctx.reset_timers(); am = C.c_int64(); amo = C.c_int64(); b = C.c_double(); nn = C.c_int64()
    L.check(lib.b7_blr_score(h, g, 0, 0.0, 0, -1.0, float(y.min()), None, C.byref(am), C.byref(amo), C.byref(b), C.byref(nn)))
    st = ctx.stage_times(); best = min(best, st["blr"][0])
out[f"blr_score_D50_M{Mb}"] = {"ms": best, "GBps": Mb * (8 * D + 16) / best * 1e-6, "TFLOPs": Mb * (D * D + 2 * D) / best * 1e-9, "cand_per_s": Mb / best * 1e3, "score_ms": st["score"][0]}
# DNGO basis on the device: 6 -> 50 -> 50 -> 50 ReLU MLP over 2^22 Sobol candidates (config 4 end to end on the GPU)
from bot7_b200 import grids, models
sob = grids.sobol({"size": Mb, "dims": 6}); gx = sob.generate_device()
Ws = [r.normal(size=(50, 6)), r.normal(size=(50, 50)) / 7, r.normal(size=(50, 50)) / 7]; bs = [np.zeros(50)] * 3
best = 1e9
for rep in range(3):
    ctx.reset_timers(); fz = models.mlp_features(gx, Ws, bs); best = min(best, ctx.stage_times()["blr"][0]); fz.free()
macs = Mb * (6 * 50 + 2 * 50 * 50)
out[f"dngo_mlp_6_50_50_50_M{Mb}"] = {"ms": best, "TFLOPs": 2 * macs / best * 1e-9, "cand_per_s": Mb / best * 1e3}
print(json.dumps(out, indent=1))
