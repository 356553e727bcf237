"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/b7_oracle.py).

The reference ships no golden vectors.  Two kinds of fixtures are written:
  * pinned.npz  -- inputs/outputs of the PINNED arithmetic (Sobol numerators, erf/cdf/pdf/EI/CB on
                   a fixed sample incl. IEEE edge cases).  The hand-traced known answers of SURVEY
                   section 4 are asserted in tests/test_oracle_golden.py independently of this file.
  * gp_*.npz    -- small GP / BLR cases of the DECLARED spec (parity unpinned: these pin the oracle
                   against drift, they are not reference outputs).
Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import b7_oracle as o  # noqa: E402


def problem(N, d, S, M, noise, seed):
    r = np.random.default_rng(seed)
    X = o.sobol_points(d, N + M)
    perm = r.permutation(N + M)
    Xo, Xc = X[np.sort(perm[:N])], X[np.sort(perm[N:])]
    y = {2: o.braninhoo, 6: o.hartmann6}.get(d, o.ackley)(Xo)
    y = (y - y.mean()) / y.std()
    hyp = np.zeros((S, d + 3))
    hyp[:, :d] = np.log(0.1) + r.random((S, d)) * (np.log(2) - np.log(0.1))
    hyp[:, d] = 0.5 * (r.random(S) - 0.5)
    hyp[:, d + 1] = 0.5 * np.log(noise)
    hyp[:, d + 2] = 0.1 * (r.random(S) - 0.5)
    return Xo, y, hyp, Xc


def main():
    r = np.random.default_rng(7)
    z = np.concatenate([r.normal(size=2000) * 3, [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-300, -1e-300, 40.0, -40.0]])
    mean = r.normal(size=4000)
    var = r.random(4000) ** 3
    var[:5] = 0.0
    mean[5] = 0.25            # sigma = 0 and imprv = 0 -> NaN
    var[5] = 0.0
    var[6] = -1.0
    mean[7] = np.nan
    fmin = 0.25
    np.savez(os.path.join(HERE, "pinned.npz"),
             sobol6=o.sobol_numerators(6, 1, 4096), sobol20_skip=o.sobol_numerators(20, 1000, 2048),
             sobol39=o.sobol_numerators(39, 1, 512), z=z, erf=o.erf_ref(z), cdf=o.norm_cdf_ref(z), pdf=o.norm_pdf_ref(z),
             mean=mean, var=var, fmin=fmin, ei=o.ei_compute(mean, var, fmin, 0.0), ei_t=o.ei_compute(mean, var, fmin, 0.1),
             lcb=o.cb_compute(mean, var, 1.0, "lower", -1.0), ucb=o.cb_compute(mean, var, 2.0, "upper", 1.0))
    for name, (N, d, S, M, noise, kern) in {"gp_c1": (50, 2, 4, 512, 1e-2, 0), "gp_h6": (200, 6, 3, 768, 1e-2, 0),
                                            "gp_m52": (130, 6, 2, 300, 1e-2, 1)}.items():
        Xo, y, hyp, Xc = problem(N, d, S, M, noise, 11)
        ei = o.acquisition(Xo, y, hyp, Xc, kern, False, o.SCORE_EI)
        cb = o.acquisition(Xo, y, hyp, Xc, kern, False, o.SCORE_CB)
        fits = [o.gp_fit(Xo, y, hyp[s], kern) for s in range(S)]
        np.savez(os.path.join(HERE, name + ".npz"), X=Xo, y=y, hyp=hyp, Xc=Xc, kernel=kern, mean=ei["mean"], var=ei["var"],
                 ei=ei["score"], ei_idx=ei["idx"], cb=cb["score"], cb_idx=cb["idx"], logml=[f["logml"] for f in fits])
    Z0 = np.maximum(r.normal(size=(300, 50)), 0)
    yb = r.normal(size=300)
    Z1 = np.maximum(r.normal(size=(700, 50)), 0)
    hb = np.array([[0.0, np.log(1e2), 0.1], [np.log(3.0), np.log(10.0), -0.2]])
    fb = [o.blr_fit(Z0, yb, h) for h in hb]
    pr = [o.blr_predict(f, Z1) for f in fb]
    np.savez(os.path.join(HERE, "blr.npz"), Z0=Z0, y=yb, Z1=Z1, hyp=hb, mean=[p[0] for p in pr], var=[p[1] for p in pr])
    print("golden written to", HERE)


if __name__ == "__main__":
    main()
