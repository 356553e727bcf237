"""CPU oracle for the bot7 surrogate-fit-and-acquisition hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product path (bot7_b200/ + libbot7_b200.so) never does.

Two kinds of content, labelled on every function:

* PINNED  -- op-for-op restatement of arithmetic that is visible in the reference tree
  (/root/reference, cited file:line).  Checked against the known answers of SURVEY.md section 4
  in tests/test_oracle_golden.py AND (round 2) against outputs of the reference source itself:
  tests/golden/make_ref_exec.py executes the reference's own Lua files, unmodified, under the Lua
  interpreter of tools/minilua and stores the results in tests/golden/ref_exec.npz;
  tests/test_ref_exec.py asserts that the functions below reproduce them bit for bit (Sobol in eight
  configurations, direction numbers, XOR, erf / cdf / pdf, EI incl. its NaN policy, the confidence
  bound, the jitter policy's iterations and factor, steal / remove, the three objectives).
* DECLARED -- the GP / Bayesian-linear-regression arithmetic lives in the un-vendored, un-versioned
  luarocks dependency "gp" (gpTorch7, bot7-scm-1.rockspec:18, models/init.lua:15).  Its source is
  not available, the reference ships no tests or golden vectors for it, and no Lua runtime exists
  in this image.  Those functions follow the textbook forms frozen in oracle/SPEC.md.
  **PARITY UNPINNED for every DECLARED function.**

numpy elementwise arithmetic rounds every operation separately (no FMA contraction), which is what
Torch7's TH elementwise loops do; LAPACK/BLAS calls go to the same family of library calls Torch7
uses (dpotrf / dtrtrs / dgemm).
"""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.linalg as sla

THREADS = int(os.environ.get("B7_ORACLE_THREADS", os.cpu_count() or 1))   # element loops; BLAS has its own pool

# --------------------------------------------------------------------------------------------
# Sobol (PINNED: grids/sobol.lua, utils/bits.lua)
# --------------------------------------------------------------------------------------------

MAX_DIMS = 40   # grids/sobol.lua:31
LOG_MAX = 30    # grids/sobol.lua:32

# grids/sobol.lua:46-52
POLY = [1, 3, 7, 11, 13, 19, 25, 37, 59, 47, 61, 55, 41, 67, 97, 91, 109, 103, 115, 131,
        193, 137, 145, 143, 241, 157, 185, 167, 229, 171, 213, 191, 253, 203, 211, 239, 247, 285, 369, 299]

# grids/sobol.lua:338-391 (create_bank): initial direction numbers, 1-based rows in the reference.
_BANK_INIT = {
    2: (3, [1, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 1,
            3, 1, 1, 3, 1, 3, 1, 3, 1, 3]),
    3: (4, [7, 5, 1, 3, 3, 7, 5, 5, 7, 7, 1, 3, 3, 7, 5, 1, 1, 5, 3, 3, 1, 7, 5, 1, 3, 3, 7,
            5, 1, 1, 5, 7, 7, 5, 1, 3, 3]),
    4: (6, [1, 7, 9, 13, 11, 1, 3, 7, 9, 5, 13, 13, 11, 3, 15, 5, 3, 15, 7, 9, 13, 9, 1, 11, 7,
            5, 15, 1, 15, 11, 5, 3, 1, 7, 9]),
    5: (8, [9, 3, 27, 15, 29, 21, 23, 19, 11, 25, 7, 13, 17, 1, 25, 29, 3, 31, 11, 5, 23, 27, 19,
            21, 5, 1, 17, 13, 7, 15, 9, 31, 9]),
    6: (14, [37, 33, 7, 5, 11, 39, 63, 27, 17, 15, 23, 29, 3, 21, 13, 31, 25, 9, 49, 33, 19, 29, 11,
             19, 27, 15, 25]),
    7: (20, [13, 33, 115, 41, 79, 17, 29, 119, 75, 73, 105, 7, 59, 65, 21, 3, 113, 61, 89, 45, 107]),
    8: (38, [7, 23, 39]),
}


def sobol_bank_unscaled(dims: int) -> np.ndarray:
    """Direction numbers m_{i,j} before scaling, rows 0..dims-1, cols 0..29.

    PINNED: create_bank (grids/sobol.lua:338-391) + the recurrence in i4_sobol
    (grids/sobol.lua:236-277): row 1 is all ones (:239); for row i with primitive polynomial
    POLY[i] of degree m, v_j = v_{j-m} XOR (2^k * v_{j-k}) over the set bits `includ[k]`.
    """
    assert 1 <= dims < MAX_DIMS
    bank = np.zeros((MAX_DIMS, LOG_MAX), dtype=np.int64)
    bank[:, 0] = 1
    for col, (first_row, vals) in _BANK_INIT.items():
        for t, v in enumerate(vals):
            bank[first_row - 1 + t, col - 1] = v
    bank[0, :] = 1                                             # :239
    for i in range(dims):                                      # :248 (starts at dimension 1)
        j, m = POLY[i] // 2, 0
        while j > 0:
            m += 1
            j //= 2
        j = POLY[i]
        includ = [0] * m
        for k in range(m - 1, -1, -1):                         # :261-265
            j2 = j // 2
            if j != 2 * j2:
                includ[k] = 1
            j = j2
        for jj in range(m, LOG_MAX):                           # :268-277
            v = int(bank[i, jj - m])
            l = 1
            for k in range(m):
                l *= 2
                if includ[k]:
                    v ^= l * int(bank[i, jj - k - 1])
            bank[i, jj] = v
    return bank[:dims].copy()


def sobol_direction_integers(dims: int) -> np.ndarray:
    """Scaled direction integers V[i][j] = m_{i,j} * 2^(30-1-j) (0-based j), uint32.

    PINNED: grids/sobol.lua:280-285 (column j multiplied by 2^(maxcol-j)), recipd = 2^-30 (:287).
    """
    bank = sobol_bank_unscaled(dims)
    scale = np.array([1 << (LOG_MAX - 1 - j) for j in range(LOG_MAX)], dtype=np.int64)
    return (bank * scale[None, :]).astype(np.uint32)


def _bit_lo0(n: int) -> int:
    """PINNED: i4_bit_lo0 (grids/sobol.lua:141-189): 1-based position of the lowest zero bit."""
    bit = 1
    i = int(n)
    i2 = i // 2
    while i != 2 * i2:
        bit += 1
        i = i2
        i2 = i // 2
    return bit


def bitwise_xor_literal(x: float, y: float, bprecis: int = 32) -> float:
    """PINNED: utils/bits.lua:27-82 -- XOR via 32-wide 0/1 vectors carried on doubles."""
    def dec2bin(dec):
        bits = [0] * bprecis
        bp = bprecis
        dec = float(dec)
        lsb = int(math.fmod(dec, 2))
        while bp > 0:
            if lsb != 0:
                bits[bp - 1] = 1
            dec = math.floor((dec - lsb) / 2)
            bp -= 1
            lsb = int(math.fmod(dec, 2))
        return bits
    bx, by = dec2bin(x), dec2bin(y)
    xor = [1 if (a + b) == 1 else 0 for a, b in zip(bx, by)]
    weights = [2.0 ** (bprecis - 1 - k) for k in range(bprecis)]
    return float(sum(w * b for w, b in zip(weights, xor)))


class SobolLiteral:
    """PINNED: literal state machine of grid:i4_sobol (grids/sobol.lua:216-335), python loops.

    Small cases only (it reproduces the reference's O(seed) catch-up loops); used to validate the
    vectorised closed form below.
    """

    def __init__(self, dims: int, literal_xor: bool = False):
        self.dims = dims
        self.V = sobol_direction_integers(dims).astype(np.int64)
        self.recipd = 2.0 ** -LOG_MAX
        self.seed = -1
        self.lastq = [0] * dims
        self.maxcol = LOG_MAX
        self._xor = (lambda a, b: int(bitwise_xor_literal(a, b))) if literal_xor else (lambda a, b: a ^ b)

    def _advance(self, l):
        for i in range(self.dims):
            self.lastq[i] = self._xor(self.lastq[i], int(self.V[i, l - 1]))

    def i4_sobol(self, seed):
        seed = max(0, int(math.floor(seed)))                   # :291
        l = None
        if seed == 0:                                           # :293
            l, self.lastq = 1, [0] * self.dims
        elif seed == self.seed + 1:                             # :295
            l = _bit_lo0(seed)
        elif seed <= self.seed:                                 # :297-305
            self.seed, l, self.lastq = 0, 1, [0] * self.dims
            for seed_temp in range(self.seed, seed):
                l = _bit_lo0(seed_temp)
                self._advance(l)
            l = _bit_lo0(seed)
        elif self.seed + 1 < seed:                              # :307-315
            for seed_temp in range(self.seed + 1, seed):
                l = _bit_lo0(seed_temp)
                self._advance(l)
            l = _bit_lo0(seed)
        if self.maxcol < l:                                     # :318-324
            return None, seed
        quasi = [q * self.recipd for q in self.lastq]           # :328-329
        self._advance(l)                                        # :330
        self.seed = seed
        return quasi, seed + 1

    def generate(self, size, skip=1, mins=None, maxes=None):
        """PINNED: grid:generate (grids/sobol.lua:58-90)."""
        g = np.full((size, self.dims), np.nan)
        for j in range(1, size + 1):
            q, _ = self.i4_sobol(j + skip - 1)
            g[j - 1] = q
        return sobol_rescale(g, mins, maxes)


def sobol_rescale(g, mins, maxes):
    """PINNED: grids/sobol.lua:79-86 (also grids/random.lua:27-33). Two separately rounded ops."""
    if mins is not None and maxes is not None:
        mins = np.asarray(mins, dtype=np.float64).reshape(1, -1)
        maxes = np.asarray(maxes, dtype=np.float64).reshape(1, -1)
        g = g * (maxes + (-mins))
        g = g + mins
    elif mins is not None:
        mins = np.asarray(mins, dtype=np.float64).reshape(1, -1)
        g = g + (mins + g.min(axis=0, keepdims=True))
    elif maxes is not None:
        maxes = np.asarray(maxes, dtype=np.float64).reshape(1, -1)
        g = g * (maxes / g.max(axis=0, keepdims=True))
    return g


def sobol_numerators(dims: int, first_seed: int, count: int) -> np.ndarray:
    """Integer numerators (x * 2^30) of sequence elements seed = first_seed .. first_seed+count-1.

    Closed form of the state machine above: lastq(seed) = XOR_{b in bits(gray(seed))} V[:, b]
    with gray(s) = s ^ (s >> 1)  (validated against SobolLiteral in tests).
    """
    V = sobol_direction_integers(dims).astype(np.uint32)
    seeds = np.arange(first_seed, first_seed + count, dtype=np.uint64)
    assert first_seed >= 0 and first_seed + count <= (1 << LOG_MAX)
    gray = seeds ^ (seeds >> np.uint64(1))
    out = np.zeros((count, dims), dtype=np.uint32)
    for b in range(LOG_MAX):
        mask = ((gray >> np.uint64(b)) & np.uint64(1)).astype(bool)
        if mask.any():
            out[mask] ^= V[None, :, b]
    return out


def sobol_points(dims, size, skip=1, mins=None, maxes=None):
    """Vectorised equivalent of SobolLiteral.generate (points j=1..size use seed=j+skip-1)."""
    num = sobol_numerators(dims, skip, size)
    g = num.astype(np.float64) * (2.0 ** -LOG_MAX)
    return sobol_rescale(g, mins, maxes)


# --------------------------------------------------------------------------------------------
# erf / normal cdf / pdf / EI / confidence bound  (PINNED: utils/math.lua, scores/*.lua)
# --------------------------------------------------------------------------------------------

SQRT2_INV = 1 / math.sqrt(2)                # utils/math.lua:13
SQRT2PI_INV = 1 / math.sqrt(2 * math.pi)    # utils/math.lua:15
_C1, _C2 = 0.254829592, -0.284496736         # utils/math.lua:263
_C3, _C4 = 1.421413741, -1.453152027         # utils/math.lua:264
_C5, _P = 1.061405429, 0.3275911             # utils/math.lua:265


def erf_ref(x):
    """PINNED: utils/math.lua:261-288 (Abramowitz-Stegun 7.1.26, every op separately rounded)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(all="ignore"):
        t = 1.0 / (np.abs(x) * _P + 1.0)                        # :280  abs, mul p, add 1, pow -1
        r = t * _C5                                             # :281
        r = r + _C4
        r = r * t
        r = r + _C3
        r = r * t                                               # :282
        r = r + _C2
        r = r * t
        r = r + _C1
        r = r * t
        e = np.exp((x * x) * -1.0)                              # :283
        r = (r * e) * -1.0 + 1.0                                # :284
        sign = (x >= 0.0).astype(np.float64) * 2.0 + -1.0       # :285
        return r * sign                                         # :286


def norm_pdf_ref(z):
    """PINNED: utils/math.lua:293-300."""
    z = np.asarray(z, dtype=np.float64)
    with np.errstate(all="ignore"):
        return np.exp((z * z) * -0.5) * SQRT2PI_INV


def norm_cdf_ref(z):
    """PINNED: utils/math.lua:305-312."""
    z = np.asarray(z, dtype=np.float64)
    with np.errstate(all="ignore"):
        return (erf_ref(z * SQRT2_INV) + 1.0) * 0.5


def ei_compute(fval, fvar, fmin, tradeoff=0.0):
    """PINNED: EI.compute (scores/expected_improvement.lua:69-88), F=1 (no fantasy mean).

    The reference reads a global `config.tradeoff` (:70); the intended value is the score's own
    config.tradeoff (default 0.0, :30) and is an explicit argument here.
    """
    fval = np.asarray(fval, dtype=np.float64)
    fvar = np.asarray(fvar, dtype=np.float64)
    with np.errstate(all="ignore"):
        sigma = np.sqrt(fvar)                                   # :73
        imprv = (fmin + (-fval)) + (-tradeoff)                  # :74
        z = imprv / sigma                                       # :75
        ei = imprv * norm_cdf_ref(z) + sigma * norm_pdf_ref(z)  # :78-79
        # :80 clamp(0, inf) -- TH clamp is comparison based: (x < lo ? lo : (x > hi ? hi : x)); NaN passes
        ei = np.where(ei < 0.0, 0.0, ei)
    return ei


def cb_compute(fval, fvar, tradeoff=1.0, bound="lower", sign=-1.0):
    """PINNED: conf_bound.compute / UCB / LCB (scores/confidence_bound.lua:70-106)."""
    fval = np.asarray(fval, dtype=np.float64)
    fvar = np.asarray(fvar, dtype=np.float64)
    with np.errstate(all="ignore"):
        s = np.sqrt(fvar) * tradeoff
        if bound.lower() == "lower":
            val = fval + (-s)                                   # :102-106
        elif bound.lower() == "upper":
            val = fval + s                                      # :96-100
        else:
            raise ValueError(bound)
    return val if sign > 0.0 else -val                          # :89-93


def mc_average(score_per_draw):
    """PINNED: bots/bayesopt.lua:69-80 -- sequential sum from 0.0 in draw order, one divide."""
    score_per_draw = np.asarray(score_per_draw, dtype=np.float64)
    acc = np.zeros(score_per_draw.shape[1], dtype=np.float64)
    for s in range(score_per_draw.shape[0]):
        acc = acc + score_per_draw[s]
    return acc / float(score_per_draw.shape[0])


def argmax_first(score):
    """PINNED: bots/bayesopt.lua:96 `score:max(1)`: first index attaining the maximum, 1-based.

    NaN policy is TH-version dependent (SURVEY a-11); the contract here: NaNs are skipped by the
    strict `>` scan, and their count is returned so the caller can decide.
    Returns (best, idx_1based, nan_count); idx 0 if every entry is NaN / the vector is empty.
    """
    score = np.asarray(score, dtype=np.float64)
    nan = np.isnan(score)
    nan_count = int(nan.sum())
    if score.size == 0 or nan_count == score.size:
        return float("nan"), 0, nan_count
    valid = np.flatnonzero(~nan)
    idx = int(valid[np.argmax(score[valid])])   # numpy argmax returns the first maximum among the non-NaN entries
    return float(score[idx]), idx + 1, nan_count


# --------------------------------------------------------------------------------------------
# jitter Cholesky (PINNED policy: utils/math.lua:159-218)
# --------------------------------------------------------------------------------------------

def potrf_lower(K):
    """LAPACK dpotrf('L'); returns (L, info) with the strict upper triangle zeroed (TH behaviour)."""
    c, info = sla.lapack.dpotrf(np.asarray(K, dtype=np.float64), lower=1, clean=1, overwrite_a=0)
    return c, int(info)


def chol_jitter(K, eps=1e-8, growth=1.1, max_eps=None):
    """PINNED: utils.math.chol retry policy (utils/math.lua:164-216).

    Returns (L, jitter_used, iterations).  First retry uses eps*growth = 1.1e-8 (:188-190: eps is
    multiplied *before* use); when eps > max_eps = ||K||_F the identity is factorised (:184-186).
    """
    K = np.asarray(K, dtype=np.float64)
    L, info = potrf_lower(K)
    if info == 0:
        return L, 0.0, 0
    if max_eps is None:
        max_eps = float(np.linalg.norm(K))
    n = K.shape[0]
    itr = 0
    while True:
        itr += 1
        # `eps > max_eps` is the reference's give-up test; a NaN or infinite norm would loop forever there
        # (documented deviation: treated as give-up, same as the CUDA path).
        if eps > max_eps or not math.isfinite(max_eps):
            L, info = potrf_lower(np.eye(n))
            return L, float("inf"), itr
        eps = eps * growth
        L, info = potrf_lower(K + eps * np.eye(n))
        if info == 0:
            return L, eps, itr


# --------------------------------------------------------------------------------------------
# GP regression  (DECLARED: oracle/SPEC.md; gpTorch7 not available -- PARITY UNPINNED)
# --------------------------------------------------------------------------------------------

KERNEL_ARDSE, KERNEL_MATERN52 = 0, 1
NOISELESS_JITTER = 1e-8      # SPEC.md: K_y = K + (sigma_n^2 + 1e-8*sigma_f^2 [noiseless]) I
LOG2PI = math.log(2 * math.pi)


def parse_hyp(hyp, d):
    """hyp row layout (SPEC.md): [log l_1..log l_d, log sigma_f, log sigma_n, m]."""
    hyp = np.asarray(hyp, dtype=np.float64).reshape(-1)
    assert hyp.size == d + 3
    w = np.exp(-hyp[:d])                       # inverse length-scales
    sf2 = math.exp(2.0 * hyp[d])
    sn2 = math.exp(2.0 * hyp[d + 1])
    return w, sf2, sn2, float(hyp[d + 2])


def _cov_block(kernel, A, B, w, sf2, out):
    """One block of the covariance (rows of A x all of B) written into `out`; the per-element operation
    order is the declared one: r2 accumulated in dimension order, every operation rounded separately."""
    r2 = np.zeros((A.shape[0], B.shape[0]))
    t = np.empty_like(r2)
    for k in range(A.shape[1]):
        np.subtract(A[:, k][:, None], B[:, k][None, :], out=t)
        np.multiply(t, w[k], out=t)
        np.multiply(t, t, out=t)
        r2 += t
    if kernel == KERNEL_ARDSE:
        np.multiply(r2, -0.5, out=r2)
        np.exp(r2, out=r2)
        np.multiply(r2, sf2, out=out)
        return out
    if kernel == KERNEL_MATERN52:
        r = np.sqrt(r2)
        s5r = math.sqrt(5.0) * r
        np.multiply(sf2, (1.0 + s5r + (5.0 / 3.0) * r2) * np.exp(-s5r), out=out)
        return out
    raise ValueError(kernel)


_COV_BLOCK_ENTRIES = 1 << 18     # block of ~2 MB per temporary: stays in a core's L2


def cov(kernel, A, B, w, sf2):
    """DECLARED. r2 = sum_d ((a_d - b_d) * w_d)^2, accumulated in d order.
    ARD-SE: sf2*exp(-r2/2); Matern-5/2: sf2*(1 + sqrt5 r + 5 r2/3) exp(-sqrt5 r).

    Evaluated in one blocked pass (row blocks of A sized for the cache, spread over the host threads like
    TH's OpenMP element loops) instead of d full M x N temporaries; every entry sees the same operations in
    the same order, so the values do not depend on the blocking."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    if kernel not in (KERNEL_ARDSE, KERNEL_MATERN52):
        raise ValueError(kernel)
    out = np.empty((A.shape[0], B.shape[0]))
    rows = max(1, _COV_BLOCK_ENTRIES // max(B.shape[0], 1))
    starts = range(0, A.shape[0], rows)
    if len(starts) <= 1 or THREADS <= 1:
        for r0 in starts:
            _cov_block(kernel, A[r0:r0 + rows], B, w, sf2, out[r0:r0 + rows])
        return out
    with ThreadPoolExecutor(THREADS) as ex:
        list(ex.map(lambda r0: _cov_block(kernel, A[r0:r0 + rows], B, w, sf2, out[r0:r0 + rows]), starts))
    return out


def gp_fit(X, y, hyp, kernel=KERNEL_ARDSE, noiseless=False):
    """DECLARED. One draw: K_y, jitter-Cholesky, alpha, log marginal likelihood."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, d = X.shape
    w, sf2, sn2, m = parse_hyp(hyp, d)
    K = cov(kernel, X, X, w, sf2)
    diag = sn2 + (NOISELESS_JITTER * sf2 if noiseless else 0.0)
    K[np.diag_indices(n)] += diag
    L, jit, iters = chol_jitter(K)
    r = y - m
    t = sla.solve_triangular(L, r, lower=True)
    alpha = sla.solve_triangular(L.T, t, lower=False)
    logml = -0.5 * float(r @ alpha) - float(np.log(np.diag(L)).sum()) - 0.5 * n * LOG2PI
    return dict(X=X, L=L, alpha=alpha, w=w, sf2=sf2, sn2=sn2, m=m, kernel=kernel,
                jitter=jit, iters=iters, logml=logml)


def gp_predict(fit, Xs, include_noise=False, chunk=8192):
    """DECLARED. mean = m + k*^T alpha; var = max(sf2 - colsumsq(L^-1 k*), 0) (+ sn2 on request).
    Candidates are processed in chunks so that K* and V (chunk x N) stay small; per candidate the arithmetic is
    the same for any chunk size up to BLAS's own blocking."""
    Xs = np.asarray(Xs, dtype=np.float64)
    M = Xs.shape[0]
    mean, var = np.empty(M), np.empty(M)
    for c0 in range(0, max(M, 1), chunk):
        Ks = cov(fit["kernel"], Xs[c0:c0 + chunk], fit["X"], fit["w"], fit["sf2"])  # chunk x N
        mean[c0:c0 + chunk] = fit["m"] + Ks @ fit["alpha"]
        V = sla.solve_triangular(fit["L"], Ks.T, lower=True)
        var[c0:c0 + chunk] = fit["sf2"] - np.einsum("ij,ij->j", V, V)
    var = np.maximum(var, 0.0)
    if include_noise:
        var = var + fit["sn2"]
    return mean, var


# --------------------------------------------------------------------------------------------
# Bayesian linear regression head used by DNGO (DECLARED; PARITY UNPINNED)
# --------------------------------------------------------------------------------------------

def blr_fit(Z0, y, hyp):
    """DECLARED. hyp = [log alpha_p, log beta, m]; A = beta Z0^T Z0 + alpha_p I; w = beta A^-1 Z0^T (y-m)."""
    Z0 = np.asarray(Z0, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    hyp = np.asarray(hyp, dtype=np.float64).reshape(-1)
    alpha_p, beta, m = math.exp(hyp[0]), math.exp(hyp[1]), float(hyp[2])
    D = Z0.shape[1]
    A = beta * (Z0.T @ Z0) + alpha_p * np.eye(D)
    L, jit, iters = chol_jitter(A)
    b = beta * (Z0.T @ (y - m))
    w = sla.solve_triangular(L.T, sla.solve_triangular(L, b, lower=True), lower=False)
    return dict(L=L, w=w, beta=beta, alpha_p=alpha_p, m=m, jitter=jit)


def blr_predict(fit, Z1):
    """DECLARED. mean = m + phi^T w ; var = ||L_A^-1 phi||^2 + 1/beta."""
    Z1 = np.asarray(Z1, dtype=np.float64)
    mean = fit["m"] + Z1 @ fit["w"]
    V = sla.solve_triangular(fit["L"], Z1.T, lower=True)
    var = np.einsum("ij,ij->j", V, V) + 1.0 / fit["beta"]
    return mean, var


def mlp_features(X, weights, biases, relu_last=True):
    """DNGO basis (models/dngo.lua:155-171): output of the last hidden layer of a Linear/ReLU stack
    (nnTools/builder.lua:133-159).  nn.Linear computes x W^T + b through BLAS (summation order unspecified),
    nn.ReLU is max(0, x): parity is to rounding, not bit-exact."""
    Z = np.asarray(X, dtype=np.float64)
    for l, (W, b) in enumerate(zip(weights, biases)):
        Z = Z @ np.asarray(W, dtype=np.float64).T + np.asarray(b, dtype=np.float64).reshape(1, -1)
        if l < len(weights) - 1 or relu_last:
            Z = np.maximum(Z, 0.0)
    return Z


# --------------------------------------------------------------------------------------------
# acquisition over a grid, marginalised over draws (PINNED control flow: bots/bayesopt.lua:56-99)
# --------------------------------------------------------------------------------------------

SCORE_EI, SCORE_CB = 0, 1


def acquisition(X, y, hyps, Xs, kernel=KERNEL_ARDSE, noiseless=False, kind=SCORE_EI, tradeoff=None,
                bound="lower", sign=-1.0, chunk=65536):
    """For each draw: fit, predict, score; sequential average; first-max argmax.
    Returns dict(score, best, idx (1-based), nan_count, mean[S,M], var[S,M])."""
    hyps = np.atleast_2d(np.asarray(hyps, dtype=np.float64))
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    Xs = np.asarray(Xs, dtype=np.float64)
    S, M = hyps.shape[0], Xs.shape[0]
    fmin = float(y.min())                                       # scores/expected_improvement.lua:64
    if tradeoff is None:
        tradeoff = 0.0 if kind == SCORE_EI else 1.0
    acc = np.zeros(M)
    means = np.empty((S, M))
    vars_ = np.empty((S, M))
    for s in range(S):
        fit = gp_fit(X, y, hyps[s], kernel, noiseless)
        for c0 in range(0, M, chunk):
            mu, var = gp_predict(fit, Xs[c0:c0 + chunk])
            means[s, c0:c0 + chunk] = mu
            vars_[s, c0:c0 + chunk] = var
        sc = ei_compute(means[s], vars_[s], fmin, tradeoff) if kind == SCORE_EI else \
            cb_compute(means[s], vars_[s], tradeoff, bound, sign)
        acc = acc + sc                                          # bots/bayesopt.lua:76
    score = acc / float(S)                                      # bots/bayesopt.lua:79
    best, idx, nan_count = argmax_first(score)
    return dict(score=score, best=best, idx=idx, nan_count=nan_count, mean=means, var=vars_)


# --------------------------------------------------------------------------------------------
# objectives used to synthesise Y_obs (PINNED: benchmarks/*.lua)
# --------------------------------------------------------------------------------------------

def braninhoo(X):
    """PINNED: benchmarks/braninhoo.lua:21-44."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    c1 = -5.1 / (4.0 * math.pi * math.pi)
    c2 = 5.0 / math.pi
    c3 = 10.0 - 10.0 / (8.0 * math.pi)
    Z1 = X[:, 0] * 15.0 + -5.0
    Z2 = X[:, 1] * 15.0
    return (((Z2 + (Z1 * Z1) * c1) + Z1 * c2) + -6.0) ** 2 + (np.cos(Z1) * c3 + 10.0)


_H6_A = -np.array([[10.0, 3.00, 17.0, 3.50, 1.70, 8.00], [0.05, 10.0, 17.0, 0.10, 8.00, 14.0],
                   [3.00, 3.50, 1.70, 10.0, 17.0, 8.00], [17.0, 8.00, 0.05, 10.0, 0.10, 14.0]])
_H6_P = -np.array([[.1312, .1696, .5569, .0124, .8283, .5886], [.2329, .4135, .8307, .3736, .1004, .9991],
                   [.2348, .1451, .3522, .2883, .3047, .6650], [.4047, .8828, .8732, .5743, .1091, .0381]])
_H6_a = -np.array([1.0, 1.2, 3.0, 3.2])


def hartmann6(X):
    """PINNED: benchmarks/hartmann6.lua:36-63: Y = a . exp(sum(A o (x + P)^2, 2)) with A, P, a negated."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    T = (X[:, None, :] + _H6_P[None, :, :]) ** 2
    E = np.exp((_H6_A[None, :, :] * T).sum(axis=2))
    return E @ _H6_a


def ackley(X):
    """PINNED: benchmarks/ackley.lua:20-49."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    a, b, c, d = 20, -0.2, 2.0 * math.pi, math.exp(1.0)
    Z = (X + -0.5) * 65.536
    t1 = np.exp(np.sqrt((Z * Z).mean(axis=1)) * b) * -a
    t2 = -np.exp(np.cos(Z * c).mean(axis=1))
    return t1 + t2 + (a + d)


OBJECTIVES = {"braninhoo": (braninhoo, 2), "hartmann6": (hartmann6, 6), "ackley": (ackley, None)}
