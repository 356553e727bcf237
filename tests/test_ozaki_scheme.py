"""The error-free int8 slicing behind the INT8 tensor path (posterior_i8.cu, potrf_i8.cu, trtri_i8.cu), restated in
exact integer arithmetic on the CPU (oracle/ozaki_ref.py) and checked against the claims written in those files."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import ozaki_ref as oz  # noqa: E402


def edge_values():
    one = np.nextafter(1.0, 0.0)
    return np.array([0.0, one, -one, 0.5, -0.5, 2.0 ** -54, -2.0 ** -54, 1.5 * 2.0 ** -55, 2.0 ** -60, 127.5 / 128, -127.5 / 128,
                     0.4980392156862745, -0.5019607843137255, 1.0 / 3.0, -2.0 / 3.0])


def test_digits_are_in_range_and_exact():
    r = np.random.default_rng(0)
    t = np.concatenate([r.uniform(-1, 1, 200000), r.uniform(-1, 1, 50000) * 2.0 ** -r.integers(0, 60, 50000), edge_values()])
    d = oz.digits(t)
    assert d[0].min() >= -64 and d[0].max() <= 64
    assert d[1:].min() >= -128 and d[1:].max() <= 127
    x = oz.undigits(d)
    assert all(int(a) == int(b) for a, b in zip(x[-15:], np.rint(edge_values() * 2.0 ** 54).astype(np.int64)))
    assert np.array_equal(x.astype(np.float64), np.rint(t * 2.0 ** 54))          # the representation IS rint(t 2^54)
    err = np.abs(x.astype(np.float64) * 2.0 ** -54 - t)
    assert err.max() <= 2.0 ** -55                                                 # half a unit of the last digit


@pytest.mark.parametrize("K", [64, 512, 4096])
def test_sliced_product_error_and_int32_headroom(K):
    r = np.random.default_rng(K)
    M, N = 24, 16
    A = r.normal(size=(M, K)) * np.exp(r.normal(size=(M, 1)) * 3)                 # rows of very different magnitude
    B = np.abs(r.normal(size=(N, K))) * np.exp(r.normal(size=(N, 1)))
    V, classes, V_full, acc, sa, sb = oz.sliced_product(A, B)
    assert np.abs(classes).max() < 2 ** 31
    assert np.abs(classes).max() <= 7 * 2 ** 14 * K                               # at most 7 products of |d e| <= 2^14 per term
    # kept part vs the exact product of the rounded operands: the dropped pairs (p + q >= 9) are below 2^-54 per term
    scale = sa[:, None] * sb[None, :]
    kept = np.array([[float(acc[i, j] * (1 << 48) - V_full[i, j]) for j in range(N)] for i in range(M)]) * 2.0 ** -108
    assert np.abs(kept).max() <= 6 * K * 2.0 ** -54
    # and against plain float64 (itself rounded): within the truncation + the rounding of the operands
    ref = A @ B.T
    assert np.abs(V - ref).max() / scale.max() <= 8 * K * 2.0 ** -54
    assert np.max(np.abs(V - ref) / scale) <= 8 * K * 2.0 ** -54


def test_worst_case_digits_fit_int32_up_to_16384_terms():
    # all digits at their extremes: the bound the kernels rely on (B7_I8_MAX_NP = 16384)
    K = 16384
    d = np.full((K,), -128, dtype=np.int64)
    s = int(d @ d)                                                                # one slice pair
    assert s == 2 ** 14 * K and 7 * s < 2 ** 31
    assert 7 * 2 ** 14 * (K + 2400) >= 2 ** 31                                    # ... and not much further


def test_power_of_two_scales_commute_with_rounding():
    # cov_slices_kernel folds sf2 / tau * 2^54 into one factor: (sf2 e) 2^k == (sf2 2^k) e bit for bit
    r = np.random.default_rng(5)
    sf2, e = np.exp(r.normal(size=1000)), r.random(1000)
    k = 54 - np.frexp(sf2)[1]
    assert np.array_equal(np.ldexp(sf2 * e, k), np.ldexp(sf2, k) * e)


@pytest.mark.parametrize("NB", list(range(1, 21)) + [32, 37])
def test_block_recursive_inverse_for_any_number_of_blocks(NB):
    # the level / pair / k-range structure of trtri_i8.cu (short or empty last pairs when NB is not a power of two)
    blk = 3
    r = np.random.default_rng(NB)
    n = NB * blk
    L = np.tril(r.normal(size=(n, n))) + 4.0 * np.eye(n)
    X = oz.blockrec_inverse(L, blk)
    assert np.allclose(X @ L, np.eye(n), atol=1e-10) and np.allclose(np.triu(X, 1), 0.0)
    # pair bookkeeping: at every level each row block belongs to exactly one half of exactly one pair
    nb = 1
    while nb < NB:
        covered = []
        for pair in range((NB + 2 * nb - 1) // (2 * nb)):
            n2 = oz.n2_of(NB, nb, pair)
            assert 0 <= n2 <= nb
            covered += list(range(pair * 2 * nb, min(NB, pair * 2 * nb + nb))) + list(range(pair * 2 * nb + nb, pair * 2 * nb + nb + n2))
        assert covered == list(range(NB))
        nb *= 2


@pytest.mark.parametrize("NB,chunk,group,n_tiles", [(32, 2, 16, 296), (32, 4, 16, 296), (1, 1, 16, 5), (5, 2, 16, 40), (18, 2, 16, 33),
                                                     (128, 2, 16, 7), (7, 1, 3, 10), (12, 32, 16, 100)])
def test_posterior_work_items_cover_every_row_block_once(NB, chunk, group, n_tiles):
    chunk = min(chunk, NB)
    items = oz.walk_items(NB, chunk, group, n_tiles)
    seen = sorted((t, rb) for t, rbs in items for rb in rbs)
    assert seen == [(t, rb) for t in range(n_tiles) for rb in range(NB)]
    # weight of an item = stages streamed = sum (rb + 1): equal for all items when the chunks pair up evenly
    w = [sum(rb + 1 for rb in rbs) for _, rbs in items]
    cpt = (NB + chunk - 1) // chunk
    if NB % chunk == 0 and cpt % 2 == 0:
        assert len(set(w)) == 1
    # tiles of at most two groups are in flight among any 148 consecutive items when a group-level fills the machine
    if len(items) >= 148 and group * ((cpt + 1) // 2) >= 128:
        for i in range(0, len(items) - 148, 37):
            assert len({t // group for t, _ in items[i:i + 148]}) <= 3


def test_sliced_posterior_of_a_small_gp_meets_the_parity_tolerance():
    # the whole INT8 posterior in exact integers on the CPU: V = L^-1 K*^T through 7 x 7 slices and 28 products (one power
    # of two tau > sf2 for K*, one per row of L^-1), var = sf2 - colsumsq(V), mean = m + V^T beta, against the oracle
    import scipy.linalg as sla
    import b7_oracle as o
    N, d, M = 300, 6, 40
    r = np.random.default_rng(8)
    X = o.sobol_points(d, N + M)
    Xo, Xc = X[:N], X[N:]
    y = o.hartmann6(Xo)
    y = (y - y.mean()) / y.std()
    hyp = np.concatenate([np.log(0.1) + r.random(d) * (np.log(2) - np.log(0.1)), [0.2, 0.5 * np.log(1e-3), 0.05]])
    fit = o.gp_fit(Xo, y, hyp, 0)
    m_ref, v_ref = o.gp_predict(fit, Xc)
    Linv = sla.solve_triangular(fit["L"], np.eye(N), lower=True)
    Ks = o.kernel_matrix(Xc, Xo, fit["w"], fit["sf2"], fit["kernel"]) if hasattr(o, "kernel_matrix") else None
    if Ks is None:                                                     # ARD-SE by hand (oracle/SPEC.md)
        D = (Xc[:, None, :] - Xo[None, :, :]) * fit["w"][None, None, :]
        Ks = fit["sf2"] * np.exp(-0.5 * np.sum(D * D, axis=2))
    tau = np.ldexp(1.0, np.frexp(fit["sf2"])[1])
    V, classes, _, _, _, _ = oz.sliced_product(Linv, Ks, sb=tau)      # V[i, c] = (L^-1 k*_c)_i
    assert np.abs(classes).max() < 2 ** 31
    beta = sla.solve_triangular(fit["L"], y - fit["m"], lower=True)
    var = fit["sf2"] - np.sum(V * V, axis=0)
    mean = fit["m"] + V.T @ beta
    assert np.max(np.abs(var - v_ref)) <= 1e-12 * fit["sf2"] and np.max(np.abs(mean - m_ref)) <= 1e-11
    assert np.max(np.abs(var - v_ref) / v_ref) <= 1e-9                # the tolerance of BASELINE.json's north star
