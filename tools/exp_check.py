"""Accuracy of the table-driven exp in cov.cu against numpy, through the public path (a 1-point GP)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bot7_b200 import models
# 1 observation at 0, l = 1, sf = 1, sn2 ~ 0, y = 1, m = 0  ->  K = 1, beta = 1, mean(x*) = k(x*, 0) = exp(-x*^2 / 2)
X = np.zeros((1, 1)); y = np.array([1.0])
hyp = np.array([[0.0, 0.0, 0.5 * np.log(1e-300), 0.0]])
f = models.GPFactors(X, y, hyp)
xs = np.sqrt(2 * np.concatenate([np.linspace(0, 700, 2000001), np.random.default_rng(0).random(1000000) * 40]))[:, None]
mu, var = f.predict(0, xs)
arg = -0.5 * (xs[:, 0] * 1.0) ** 2
ref = np.exp(arg)
ok = ref > 1e-300
ulp = np.abs(mu[ok] - ref[ok]) / np.spacing(ref[ok])
print("max ulp err", ulp.max(), "mean", ulp.mean(), "frac exact", np.mean(ulp == 0), "max rel", np.max(np.abs(mu[ok] - ref[ok]) / ref[ok]))
print("tiny:", mu[~ok][:3], ref[~ok][:3])
