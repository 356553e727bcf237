"""The oracle against the known answers of SURVEY.md section 4 (hand-traced from the reference
source: grids/sobol.lua, utils/math.lua, scores/expected_improvement.lua) and the committed golden
fixtures.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_sobol_first_points(oracle):
    p = oracle.sobol_points(2, 5)
    assert p.tolist() == [[0.5, 0.5], [0.75, 0.25], [0.25, 0.75], [0.375, 0.375], [0.875, 0.875]]
    p6 = oracle.sobol_points(6, 5)
    assert p6[3].tolist() == [0.375, 0.375, 0.625, 0.125, 0.875, 0.875]
    assert p6[4].tolist() == [0.875, 0.875, 0.125, 0.625, 0.375, 0.375]


def test_sobol_sha256_65536x6(oracle):
    num = oracle.sobol_numerators(6, 1, 65536)
    assert hashlib.sha256(num.astype("<u4").tobytes()).hexdigest() == \
        "6ed6741e43f7e1a738ef44a3ed0db34e02aee5adfa796f2ff8f7f8140f4d81a1"
    assert num[-1].tolist() == [24576, 536862720, 908058624, 136421376, 949952512, 275144704]


def test_direction_numbers(oracle):
    assert oracle.sobol_bank_unscaled(7)[6, :8].tolist() == [1, 1, 3, 7, 31, 47, 109, 173]
    assert oracle.sobol_direction_integers(20)[19, :6].tolist() == \
        [536870912, 805306368, 134217728, 1006632960, 570425344, 1056964608]


@pytest.mark.parametrize("dims,size,skip", [(6, 300, 1), (20, 200, 1), (3, 64, 5), (39, 40, 1), (2, 50, 0)])
def test_closed_form_matches_literal_state_machine(oracle, dims, size, skip):
    lit = oracle.SobolLiteral(dims).generate(size, skip)
    assert np.array_equal(lit, oracle.sobol_points(dims, size, skip))


def test_literal_xor_emulation(oracle):
    # utils/bits.lua XOR-on-doubles equals integer XOR for operands < 2^30
    r = np.random.default_rng(0)
    for a, b in r.integers(0, 2 ** 30, size=(50, 2)):
        assert oracle.bitwise_xor_literal(float(a), float(b)) == float(int(a) ^ int(b))
    lit = oracle.SobolLiteral(3, literal_xor=True).generate(33, 1)
    assert np.array_equal(lit, oracle.sobol_points(3, 33, 1))


def test_literal_generator_restart_branches(oracle):
    # seed <= self.seed (restart) and skip-ahead branches of i4_sobol (grids/sobol.lua:297-315)
    g = oracle.SobolLiteral(4)
    a = g.generate(20, 1)
    b = g.generate(20, 1)          # second call restarts from seed 1 <= self.seed
    c = g.generate(5, 40)          # jump ahead
    assert np.array_equal(a, b)
    assert np.array_equal(c, oracle.sobol_points(4, 5, 40))


def test_sobol_rescale_two_rounded_ops(oracle):
    mins, maxes = np.array([-1.0, 0.1, 3.0]), np.array([2.0, 0.7, 3.5])
    g = oracle.sobol_points(3, 100)
    want = g * (maxes - mins)[None, :] + mins[None, :]
    assert np.array_equal(oracle.sobol_points(3, 100, 1, mins, maxes), want)


def test_erf_cdf_pdf_ei_known_answers(oracle):
    assert float(oracle.erf_ref(0.5)) == 0.5205000163047472
    assert float(oracle.erf_ref(0.0)) == 9.999999717180685e-10       # not 0: sign = (x>=0)*2-1
    assert float(oracle.norm_cdf_ref(-1.0)) == 0.15865526383236372
    assert float(oracle.norm_pdf_ref(1.0)) == 0.24197072451914337
    assert float(oracle.ei_compute(0.3, 0.04, 0.1, 0.0)) == 0.016663092137355943


def test_erf_accuracy_envelope(oracle):
    import scipy.special as sp
    x = np.linspace(-6, 6, 20001)
    assert np.max(np.abs(oracle.erf_ref(x) - sp.erf(x))) < 1.5e-7     # A&S 7.1.26 bound


def test_ei_ieee_edge_cases(oracle):
    # sigma = 0: z = +-inf -> ei = max(imprv, 0); sigma = 0 and imprv = 0: NaN; var < 0: NaN
    assert float(oracle.ei_compute(0.0, 0.0, 1.0)) == 1.0
    assert float(oracle.ei_compute(2.0, 0.0, 1.0)) == 0.0
    assert np.isnan(oracle.ei_compute(1.0, 0.0, 1.0))
    assert np.isnan(oracle.ei_compute(0.0, -1.0, 1.0))
    assert np.isnan(oracle.ei_compute(np.nan, 1.0, 1.0))


def test_confidence_bound_defaults(oracle):
    # defaults: tradeoff 1, bound lower, sign -1 -> -(mu - sqrt(var))
    assert float(oracle.cb_compute(0.5, 0.25)) == -(0.5 - 0.5)
    assert float(oracle.cb_compute(0.5, 0.25, 2.0, "upper", 1.0)) == 1.5
    assert float(oracle.cb_compute(0.5, 0.25, 2.0, "upper", -1.0)) == -1.5


def test_mc_average_is_sequential(oracle):
    s = np.array([[1e16], [1.0], [-1e16], [1.0]])
    assert oracle.mc_average(s)[0] == ((((0.0 + 1e16) + 1.0) - 1e16) + 1.0) / 4.0


def test_argmax_first_rule(oracle):
    assert oracle.argmax_first([0.0, 2.0, 2.0, 1.0])[1] == 2
    assert oracle.argmax_first([0.0, 0.0, 0.0])[1] == 1
    b, i, n = oracle.argmax_first([np.nan, 1.0, np.nan, 3.0])
    assert (b, i, n) == (3.0, 4, 2)
    assert oracle.argmax_first([np.nan, np.nan])[1:] == (0, 2)
    assert oracle.argmax_first([])[1] == 0
    assert oracle.argmax_first([-np.inf, -np.inf])[1] == 1


def test_jitter_policy(oracle):
    # utils/math.lua:168-216: first retry 1.1e-8, growth 1.1, gives up at eps > ||K||_F -> chol(I)
    K = np.ones((4, 4))
    L, eps, itr = oracle.chol_jitter(K)
    assert itr >= 1 and eps == pytest.approx(1e-8 * 1.1 ** itr)
    assert np.allclose(L @ L.T, K + eps * np.eye(4))
    L, eps, itr = oracle.chol_jitter(np.eye(3))
    assert (eps, itr) == (0.0, 0)
    L, eps, itr = oracle.chol_jitter(-np.eye(3))          # PD once eps > 1 (< ||K||_F = sqrt(3))
    assert 1.0 < eps < 1.1 ** 2 and itr == int(np.ceil(np.log(1e8) / np.log(1.1)))
    L, eps, itr = oracle.chol_jitter(-np.eye(3), max_eps=0.5)   # explicit cap (config.max_eps)
    assert np.isinf(eps) and np.array_equal(L, np.eye(3))


def test_benchmark_objectives(oracle):
    # published optima (benchmarks/hartmann6.lua header; Branin minimum 0.397887)
    x6 = np.array([[.201690, .150011, .476874, .275332, .311652, .657300]])
    assert oracle.hartmann6(x6)[0] == pytest.approx(-3.32237, abs=1e-5)
    xb = np.array([[(np.pi + 5) / 15, 2.275 / 15]])
    assert oracle.braninhoo(xb)[0] == pytest.approx(0.397887, abs=1e-6)
    assert oracle.ackley(np.full((1, 20), 0.5))[0] == pytest.approx(0.0, abs=1e-12)


def test_gp_spec_identities(oracle):
    r = np.random.default_rng(3)
    X, y = r.random((40, 3)), r.normal(size=40)
    hyp = np.array([np.log(0.3), np.log(0.5), np.log(0.8), 0.1, 0.5 * np.log(1e-2), 0.05])
    for kern in (0, 1):
        f = oracle.gp_fit(X, y, hyp, kern)
        K = oracle.cov(kern, X, X, f["w"], f["sf2"]) + f["sn2"] * np.eye(40)
        assert np.allclose(f["L"] @ f["L"].T, K)
        assert np.allclose(K @ f["alpha"], y - f["m"])
        mu, var = oracle.gp_predict(f, X)
        # at the observations the latent posterior is K(K+sn2 I)^-1 (y-m)
        assert np.allclose(mu, f["m"] + (K - f["sn2"] * np.eye(40)) @ f["alpha"])
        assert (var >= 0).all() and (var <= f["sf2"] + 1e-12).all()
        import scipy.stats as st
        ref = st.multivariate_normal(mean=np.full(40, f["m"]), cov=K).logpdf(y)
        assert f["logml"] == pytest.approx(ref, rel=1e-10)


def test_golden_fixtures_match_oracle(oracle):
    g = np.load(os.path.join(GOLD, "pinned.npz"))
    assert np.array_equal(g["sobol6"], oracle.sobol_numerators(6, 1, 4096))
    assert np.array_equal(g["sobol20_skip"], oracle.sobol_numerators(20, 1000, 2048))
    assert np.array_equal(g["erf"], oracle.erf_ref(g["z"]), equal_nan=True)
    assert np.array_equal(g["ei"], oracle.ei_compute(g["mean"], g["var"], float(g["fmin"]), 0.0), equal_nan=True)
    assert np.array_equal(g["lcb"], oracle.cb_compute(g["mean"], g["var"], 1.0, "lower", -1.0), equal_nan=True)
    c = np.load(os.path.join(GOLD, "gp_c1.npz"))
    a = oracle.acquisition(c["X"], c["y"], c["hyp"], c["Xc"], int(c["kernel"]), False, oracle.SCORE_EI)
    assert a["idx"] == int(c["ei_idx"])
    assert np.allclose(a["score"], c["ei"], rtol=1e-12, atol=1e-300)
