"""The DECLARED part of the oracle (oracle/SPEC.md: GP regression and the BLR head, whose reference arithmetic
lives in the absent gpTorch7 rock) against independent evaluations of the same declared forms:

* hand-derived closed forms of the N = 1 and N = 2 posteriors and log marginal likelihoods;
* 50-digit mpmath evaluations (kernel, Cholesky-free solves through mp.lu_solve, log-determinant) at small N;
* the error of the explicit-inverse formulation the CUDA path uses (V = L^-1 K*^T with L^-1 formed explicitly) next
  to the triangular-solve formulation, both measured against the mpmath reference (note ii).

This pins the arithmetic error of the declared spec.  It cannot pin the spec itself against gpTorch7 (parity stays
"unpinned": the rock's source, tests and golden vectors are not available).  CPU only."""
import math

import mpmath as mp
import numpy as np
import scipy.linalg as sla

mp.mp.dps = 50


def mp_gp(X, y, hyp, kernel, noiseless, Xs):
    """Declared forms of oracle/SPEC.md evaluated in 50-digit arithmetic from the fp64 inputs."""
    n, d = X.shape
    w = [mp.e ** (-mp.mpf(float(h))) for h in hyp[:d]]
    sf2 = mp.e ** (2 * mp.mpf(float(hyp[d])))
    sn2 = mp.e ** (2 * mp.mpf(float(hyp[d + 1])))
    m = mp.mpf(float(hyp[d + 2]))

    def k(a, b):
        r2 = mp.mpf(0)
        for i in range(d):
            t = (mp.mpf(float(a[i])) - mp.mpf(float(b[i]))) * w[i]
            r2 += t * t
        if kernel == 0:
            return sf2 * mp.e ** (-r2 / 2)
        r = mp.sqrt(r2)
        return sf2 * (1 + mp.sqrt(5) * r + 5 * r2 / 3) * mp.e ** (-mp.sqrt(5) * r)

    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = k(X[i], X[j])
        K[i, i] += sn2 + (mp.mpf("1e-8") * sf2 if noiseless else 0)
    r = mp.matrix([mp.mpf(float(v)) - m for v in y])
    alpha = mp.lu_solve(K, r)
    logml = -(r.T * alpha)[0] / 2 - mp.log(mp.det(K)) / 2 - mp.mpf(n) / 2 * mp.log(2 * mp.pi)
    means, variances = [], []
    for xs in Xs:
        ks = mp.matrix([k(xs, X[i]) for i in range(n)])
        means.append(m + (ks.T * alpha)[0])
        variances.append(sf2 - (ks.T * mp.lu_solve(K, ks))[0])
    return [float(v) for v in means], [float(v) for v in variances], float(logml), float(sf2)


def test_n1_closed_form(oracle):
    x, yv = np.array([[0.3, 0.7]]), np.array([1.25])
    hyp = np.array([math.log(0.4), math.log(0.9), 0.2, 0.5 * math.log(0.05), -0.1])
    Xs = np.array([[0.3, 0.7], [0.5, 0.1], [5.0, 5.0]])
    fit = oracle.gp_fit(x, yv, hyp, 0)
    mu, var = oracle.gp_predict(fit, Xs)
    sf2, sn2, m = math.exp(0.4), 0.05, -0.1
    for i, xs in enumerate(Xs):
        r2 = ((xs[0] - 0.3) / 0.4) ** 2 + ((xs[1] - 0.7) / 0.9) ** 2
        ks = sf2 * math.exp(-0.5 * r2)
        assert abs(mu[i] - (m + ks * (1.25 - m) / (sf2 + sn2))) <= 1e-15 * max(1.0, abs(mu[i]))
        assert abs(var[i] - (sf2 - ks * ks / (sf2 + sn2))) <= 4e-16 * sf2
    want = -0.5 * (1.25 - m) ** 2 / (sf2 + sn2) - 0.5 * math.log(sf2 + sn2) - 0.5 * math.log(2 * math.pi)
    assert abs(fit["logml"] - want) <= 1e-15 * abs(want)
    # far away the posterior is the prior: mean m, latent variance sf2 (noise not included: SPEC.md)
    assert abs(mu[2] - m) <= 1e-15 and abs(var[2] - sf2) <= 1e-15


def test_n2_closed_form(oracle):
    X, y = np.array([[0.1], [0.6]]), np.array([0.4, -0.9])
    hyp = np.array([math.log(0.35), 0.1, 0.5 * math.log(0.02), 0.05])
    ell, sf2, sn2, m = 0.35, math.exp(0.2), 0.02, 0.05
    kf = lambda a, b: sf2 * math.exp(-0.5 * ((a - b) / ell) ** 2)
    a, b = sf2 + sn2, kf(0.1, 0.6)
    det = a * a - b * b
    r0, r1 = y[0] - m, y[1] - m
    al0, al1 = (a * r0 - b * r1) / det, (a * r1 - b * r0) / det
    fit = oracle.gp_fit(X, y, hyp, 0)
    Xs = np.array([[0.0], [0.35], [0.6], [1.0]])
    mu, var = oracle.gp_predict(fit, Xs)
    for i, xs in enumerate(Xs[:, 0]):
        k0, k1 = kf(xs, 0.1), kf(xs, 0.6)
        assert abs(mu[i] - (m + k0 * al0 + k1 * al1)) <= 1e-14
        assert abs(var[i] - (sf2 - (a * k0 * k0 - 2 * b * k0 * k1 + a * k1 * k1) / det)) <= 1e-14 * sf2
    want = -0.5 * (r0 * al0 + r1 * al1) - 0.5 * math.log(det) - math.log(2 * math.pi)
    assert abs(fit["logml"] - want) <= 1e-14 * abs(want)


def test_gp_against_mpmath(oracle):
    r = np.random.default_rng(11)
    for kernel, noiseless, noise in ((0, False, 1e-2), (1, False, 1e-2), (0, True, 1e-6)):
        n, d = 24, 3
        X, Xs = r.random((n, d)), r.random((6, d))
        y = np.sin(4 * X.sum(1)) + 0.1 * r.normal(size=n)
        hyp = np.array([math.log(0.4), math.log(0.6), math.log(0.3), 0.1, 0.5 * math.log(noise), 0.05])
        fit = oracle.gp_fit(X, y, hyp, kernel, noiseless)
        mu, var = oracle.gp_predict(fit, Xs)
        mu_x, var_x, logml_x, sf2 = mp_gp(X, y, hyp, kernel, noiseless, Xs)
        # cond(K) ~ 1e3 (noise 1e-2) .. 1e7 (noiseless): errors scale with cond * eps
        tol = 1e-12 if not noiseless else 1e-8
        assert np.max(np.abs(mu - mu_x)) <= tol
        assert np.max(np.abs(np.maximum(var_x, 0.0) - var)) <= tol * sf2
        assert abs(fit["logml"] - logml_x) <= 1e-11 * abs(logml_x)


def test_explicit_inverse_against_triangular_solve_and_mpmath(oracle):
    """the CUDA path multiplies by an explicitly formed L^-1 (dtrtri order) instead of solving with
    L.  Both formulations against the 50-digit reference: the explicit inverse stays within a small factor of dtrtrs and
    orders of magnitude inside the 1e-9 bar."""
    r = np.random.default_rng(5)
    n, d = 40, 2
    X, Xs = r.random((n, d)), r.random((8, d))
    y = np.cos(3 * X[:, 0]) * X[:, 1]
    hyp = np.array([math.log(0.5), math.log(0.7), 0.0, 0.5 * math.log(1e-2), 0.0])
    fit = oracle.gp_fit(X, y, hyp, 0)
    _, var_x, _, sf2 = mp_gp(X, y, hyp, 0, False, Xs)
    Ks = oracle.cov(0, Xs, X, fit["w"], fit["sf2"])
    V_solve = sla.solve_triangular(fit["L"], Ks.T, lower=True)
    Linv, info = sla.lapack.dtrtri(fit["L"], lower=1)
    assert info == 0
    V_inv = np.tril(Linv) @ Ks.T
    e_solve = np.max(np.abs((sf2 - (V_solve ** 2).sum(0)) - var_x)) / sf2
    e_inv = np.max(np.abs((sf2 - (V_inv ** 2).sum(0)) - var_x)) / sf2
    assert e_solve <= 1e-13 and e_inv <= 1e-12
    assert e_inv <= 20 * max(e_solve, 2e-16)


def test_blr_against_mpmath(oracle):
    r = np.random.default_rng(2)
    n, D = 60, 5
    Z0, Z1 = np.maximum(r.normal(size=(n, D)), 0), np.maximum(r.normal(size=(7, D)), 0)
    y = Z0 @ r.normal(size=D) + 0.1 * r.normal(size=n)
    hyp = np.array([math.log(1.3), math.log(80.0), 0.02])
    fit = oracle.blr_fit(Z0, y, hyp)
    mu, var = oracle.blr_predict(fit, Z1)
    alpha_p, beta, m = mp.e ** mp.mpf(float(hyp[0])), mp.e ** mp.mpf(float(hyp[1])), mp.mpf(float(hyp[2]))
    Zm = mp.matrix(Z0.tolist())
    A = beta * (Zm.T * Zm) + alpha_p * mp.eye(D)
    w = mp.lu_solve(A, beta * (Zm.T * mp.matrix([mp.mpf(float(v)) - m for v in y])))
    for i in range(Z1.shape[0]):
        phi = mp.matrix(Z1[i].tolist())
        assert abs(mu[i] - float(m + (phi.T * w)[0])) <= 1e-12
        assert abs(var[i] - float((phi.T * mp.lu_solve(A, phi))[0] + 1 / beta)) <= 1e-13
