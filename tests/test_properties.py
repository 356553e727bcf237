"""Property tests (hypothesis): Sobol index-range independence, tie-break / NaN rules of the argmax and of
its sharded combine.  CPU versions run against the oracle; the `gpu` ones drive the CUDA kernels."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from bot7_b200 import parallel


@settings(max_examples=40, deadline=None)
@given(dims=st.integers(1, 39), skip=st.integers(0, 600), size=st.integers(1, 60))
def test_sobol_closed_form_equals_state_machine(oracle, dims, skip, size):
    assert np.array_equal(oracle.SobolLiteral(dims).generate(size, skip), oracle.sobol_points(dims, size, skip))


@settings(max_examples=40, deadline=None)
@given(dims=st.integers(1, 39), first=st.integers(0, (1 << 30) - 5000), n=st.integers(1, 400), cut=st.integers(0, 399))
def test_sobol_ranges_are_independent(oracle, dims, first, n, cut):
    cut = min(cut, n - 1)
    whole = oracle.sobol_numerators(dims, first, n)
    assert np.array_equal(whole[cut:], oracle.sobol_numerators(dims, first + cut, n - cut))
    assert whole.max() < (1 << 30)


scores_st = st.lists(st.one_of(st.sampled_from([0.0, 1.0, -1.0, 2.5, float("inf"), float("-inf"), float("nan")]),
                               st.floats(-10, 10, allow_nan=False)), min_size=0, max_size=60)


@settings(max_examples=150, deadline=None)
@given(vals=scores_st)
def test_argmax_first_matches_a_strict_greater_scan(oracle, vals):
    best, idx, nans = oracle.argmax_first(vals)
    b, i, n = None, 0, 0
    for k, v in enumerate(vals):          # TH-style scan: strict >, NaNs skipped and counted
        if v != v:
            n += 1
        elif b is None or v > b:
            b, i = v, k + 1
    assert (idx, nans) == (i, n)
    assert (best == b) or (b is None and best != best)


@settings(max_examples=150, deadline=None)
@given(vals=scores_st, world=st.integers(1, 8))
def test_sharded_combine_equals_global_argmax(oracle, vals, world):
    trips = []
    for g in range(world):
        r0, cnt = parallel.shard_range(len(vals), world, g)
        b, i, n = oracle.argmax_first(vals[r0:r0 + cnt])
        trips.append((b, i + r0 if i else 0, n))
    gb, gi, gn = oracle.argmax_first(vals)
    b, i, n = parallel.combine_argmax(trips)
    assert (i, n) == (gi, gn) and ((b == gb) or (gi == 0 and b != b))


@pytest.mark.gpu
@settings(max_examples=25, deadline=None)
@given(dims=st.integers(1, 39), first=st.integers(0, (1 << 30) - 20000), n=st.integers(1, 9000), affine=st.booleans())
def test_gpu_sobol_any_range(ctx, oracle, dims, first, n, affine):
    from bot7_b200 import _lib as L
    mins = np.linspace(-2.0, 1.0, dims) if affine else None
    maxes = np.linspace(1.5, 7.0, dims) if affine else None
    out = np.empty((n, dims))
    L.check(L.lib().b7_sobol_generate(ctx.handle, dims, first, n, L.dptr(mins), L.dptr(maxes), L.dptr(out), None))
    ref = oracle.sobol_rescale(oracle.sobol_numerators(dims, first, n).astype(np.float64) * 2.0 ** -30, mins, maxes)
    assert np.array_equal(out, ref)


@pytest.mark.gpu
@settings(max_examples=40, deadline=None)
@given(vals=st.lists(st.sampled_from([0.0, 0.25, 1.0, 4.0, float("nan")]), min_size=1, max_size=700), S=st.integers(1, 3))
def test_gpu_argmax_ties_and_nans(ctx, oracle, vals, S):
    # confidence bound with tradeoff 0, sign +1 returns the mean itself: scores are exactly `vals` (averaged over S
    # identical draws), so ties and NaNs land where we put them
    from bot7_b200 import _lib as L
    M = len(vals)
    mean = np.tile(np.array(vals), (S, 1))
    var = np.ones((S, M))
    sc = np.empty(M)
    am, best, nn = C.c_int64(), C.c_double(), C.c_int64()
    L.check(L.lib().b7_score_moments(ctx.handle, L.SCORE_CB, L.dptr(mean), L.dptr(var), S, M, 0.0, L.BOUND_UPPER, 1.0, 0.0,
                                     L.dptr(sc), C.byref(am), C.byref(best), C.byref(nn)))
    ref = oracle.mc_average([oracle.cb_compute(mean[s], var[s], 0.0, "upper", 1.0) for s in range(S)])
    assert np.array_equal(sc, ref, equal_nan=True)
    b, i, n = oracle.argmax_first(ref)
    assert (am.value, nn.value) == (i, n)
