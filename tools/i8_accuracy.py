"""Deviation of the INT8 (error-free sliced) posterior from the FP64 DMMA posterior on the same factors (development aid)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b7_oracle as o
from bot7_b200 import _lib as L, models

ctx = L.Context.default()
out = []
for (N, d, M, kern, noise) in ((900, 6, 8192, "ardse", 1e-2), (4096, 6, 16384, "ardse", 1e-2), (4096, 6, 16384, "matern52", 1e-4), (2048, 20, 8192, "ardse", 1e-6)):
    r = np.random.default_rng(N + d)
    X = o.sobol_points(d, N + M); Xo, Xc = X[:N].copy(), X[N:].copy()
    y = (o.hartmann6(Xo) if d == 6 else o.ackley(Xo)); y = (y - y.mean()) / y.std()
    S = 2
    hyp = np.zeros((S, d + 3)); hyp[:, :d] = np.log(0.1) + r.random((S, d)) * (np.log(2) - np.log(0.1))
    hyp[:, d] = 0.5 * (r.random(S) - 0.5); hyp[:, d + 1] = 0.5 * np.log(noise); hyp[:, d + 2] = 0.1 * (r.random(S) - 0.5)
    f = models.GPFactors(Xo, y, hyp, kern)
    res = {}
    for path in (L.PATH_FP64_DMMA, L.PATH_INT8_OZAKI):
        ctx.set_posterior_path(path)
        res[path] = [f.predict(s, Xc) for s in range(S)]
    dv = max(np.max(np.abs(res[0][s][1] - res[1][s][1])) / np.exp(2 * hyp[s, d]) for s in range(S))
    dm = max(np.max(np.abs(res[0][s][0] - res[1][s][0])) for s in range(S))
    vmin = min(res[0][s][1].min() / np.exp(2 * hyp[s, d]) for s in range(S))
    out.append({"N": N, "d": d, "M": M, "kernel": kern, "noise": noise, "max_var_diff_over_sf2": dv, "max_mean_diff": dm, "min_var_over_sf2": vmin})
    f.free()
ctx.set_posterior_path(L.PATH_INT8_OZAKI)
print(json.dumps(out))
