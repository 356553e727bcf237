"""Density-evaluation cost vs batch width (how wide should the speculative sampler go?)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b7_oracle as o
from bot7_b200 import _lib as L, models
out = {}
for N in (128, 512, 2048, 4096):
    d = 6; X = o.sobol_points(d, N); y = o.hartmann6(X); y = (y - y.mean()) / y.std()
    for S in (1, 2, 4, 8, 16):
        h = np.zeros((S, d + 3)); h[:, :d] = np.log(0.4); h[:, d + 1] = 0.5 * np.log(1e-2)
        f = models.GPFactors(X, y, h, flags=L.FIT_LOGML_ONLY)
        for _ in range(2): f.refit(h + 0.01, L.FIT_LOGML_ONLY)
        t0 = time.perf_counter()
        for i in range(8): f.refit(h + 0.001 * i, L.FIT_LOGML_ONLY)
        out[f"N{N}_S{S}"] = round((time.perf_counter() - t0) / 8 * 1e3, 3)
        f.free()
print(json.dumps(out))
