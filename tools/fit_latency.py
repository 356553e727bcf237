"""Single-factor density evaluation latency (the slice sampler's f): fresh fit vs resident refit."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b7_oracle as o
from bot7_b200 import _lib as L, models
out = {}
for N in (50, 512, 2048, 4096, 8192):
    d = 6
    X = o.sobol_points(d, N); y = o.hartmann6(X); y = (y - y.mean()) / y.std()
    h = np.zeros((1, d + 3)); h[0, :d] = np.log(0.4); h[0, d + 1] = 0.5 * np.log(1e-2)
    f = models.GPFactors(X, y, h, flags=L.FIT_LOGML_ONLY)
    for _ in range(3): f.refit(h + 0.01, L.FIT_LOGML_ONLY)
    t0 = time.perf_counter()
    for i in range(10): f.refit(h + 0.001 * i, L.FIT_LOGML_ONLY)
    t_refit = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for i in range(5):
        g = models.GPFactors(X, y, h, flags=L.FIT_LOGML_ONLY); g.free()
    t_fresh = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter(); ref = o.gp_fit(X, y, h[0], 0); t_cpu = time.perf_counter() - t0
    out[f"N{N}"] = {"refit_ms": t_refit * 1e3, "fresh_fit_ms": t_fresh * 1e3, "cpu_oracle_ms": t_cpu * 1e3,
                    "logml_rel_err": abs(f.refit(h, L.FIT_LOGML_ONLY).logml[0] - ref["logml"]) / abs(ref["logml"])}
    f.free()
print(json.dumps(out, indent=1))
