"""Host-side logic that needs no GPU: sharding, deterministic argmax combine (also across two
gloo ranks), slice-sampler control flow, config defaults."""
import os
import sys

import numpy as np
import pytest

from bot7_b200 import parallel, samplers, scores

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("M,G", [(10, 3), (8, 8), (5, 8), (1 << 20, 7), (0, 4)])
def test_shard_ranges_partition(M, G):
    rows = [parallel.shard_range(M, G, g) for g in range(G)]
    assert rows[0][0] == 0 and sum(c for _, c in rows) == M
    for (a, ca), (b, _) in zip(rows, rows[1:]):
        assert a + ca == b
    assert max(c for _, c in rows) - min(c for _, c in rows) <= 1


def test_combine_argmax_rule():
    # max score; ties -> smallest global index (TH first-max, bots/bayesopt.lua:96); empty shards skipped
    assert parallel.combine_argmax([(1.0, 9, 0), (2.0, 40, 1), (2.0, 17, 0)]) == (2.0, 17, 1)
    assert parallel.combine_argmax([(float("nan"), 0, 3), (0.0, 5, 0)]) == (0.0, 5, 3)
    b, i, n = parallel.combine_argmax([(float("nan"), 0, 2)])
    assert i == 0 and n == 2 and np.isnan(b)
    assert parallel.combine_argmax([(-np.inf, 3, 0), (-np.inf, 2, 0)])[1] == 2


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from bot7_b200 import parallel as par
    # each rank scores its own shard of a known score vector
    score = np.array([0.1, 0.7, 0.7, 0.2, np.nan, 0.7, 0.3])
    r0, cnt = par.shard_range(score.size, world, rank)
    loc = score[r0:r0 + cnt]
    ok = ~np.isnan(loc)
    if ok.any():
        j = int(np.argmax(np.where(ok, loc, -np.inf)))
        trip = (float(loc[j]), r0 + j + 1, int((~ok).sum()))
    else:
        trip = (float("nan"), 0, int((~ok).sum()))
    q.put((rank, par.allgather_argmax(*trip)))
    dist.destroy_process_group()


def _gloo_gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from bot7_b200 import parallel as par
    ok = True
    for S, per_draw in ((4, 6), (5, 3), (1, 4)):          # even split, uneven split, fewer draws than ranks
        full = torch.full((S * per_draw,), -1.0, dtype=torch.float64)
        s0, cnt = par.draw_range(S, world, rank)
        for s in range(s0, s0 + cnt):                     # "factorise" my draws
            full[s * per_draw:(s + 1) * per_draw] = torch.arange(per_draw, dtype=torch.float64) + 100.0 * s
        par.allgather_draws(full, per_draw, S, world, rank)
        want = torch.cat([torch.arange(per_draw, dtype=torch.float64) + 100.0 * s for s in range(S)])
        ok = ok and bool(torch.equal(full, want))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_draw_sharded_allgather_two_gloo_ranks():
    # the fit exchange of the multi-GPU path (each rank factorises S/G draws, then all-gathers them in place)
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctxm.Process(target=_gloo_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == {0: True, 1: True}


def test_allgather_argmax_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctxm.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res[0] == res[1] == (0.7, 2, 1)      # first maximum globally, NaN counted once


def test_slice_sampler_control_flow():
    # log-density of N(1, 0.5^2) in 2-D; the sampler must leave it invariant
    rng = np.random.default_rng(0)
    calls = []

    def f(x, _):
        calls.append(1)
        return float(-0.5 * np.sum(((x - 1.0) / 0.5) ** 2))

    s = samplers.slice()
    x = np.zeros((1, 2))
    out = []
    for _ in range(600):
        x = s(f, x, {"nSamples": 1}, None, rng=rng)
        out.append(x[0])
    out = np.array(out[100:])
    assert abs(out.mean() - 1.0) < 0.1 and abs(out.std() - 0.5) < 0.1
    assert 4 <= len(calls) / 600 <= 40          # SURVEY a-15: typically 5-30 density evaluations per sample
    opt = samplers.slice.configure({"step_out": False})
    assert opt["step_out"] is False and opt["logspace"] is True and opt["max_step"] == 1e3 and opt["nSamples"] == 1
    g = s(f, np.zeros((1, 2)), {"nSamples": 3, "gibbs": True}, None, rng=rng)
    assert g.shape == (3, 2)


def test_speculative_sampler_reproduces_the_sequential_chain():
    # batched / speculative density evaluations must not change a single bit of the chain nor the RNG state
    def f(x, _):
        return float(-0.5 * np.sum(((x - 1.0) / np.array([0.5, 2.0, 0.1])) ** 2))

    def fb(X, _):
        return [f(X[i:i + 1], None) for i in range(X.shape[0])]

    for seed, opt in ((0, {}), (1, {"width": 0.37}), (2, {"step_out": False}), (3, {"widths": np.array([[0.3, 2.5, 0.05]])})):
        r1, r2 = np.random.default_rng(seed), np.random.default_rng(seed)
        x1 = x2 = np.zeros((1, 3))
        seq_evals, calls = [0], 0

        def fc(x, a):
            seq_evals[0] += 1
            return f(x, a)
        sp = samplers.slice_speculative()
        for _ in range(40):
            x1 = samplers.slice()(fc, x1, dict(opt), None, rng=r1)
            x2 = sp(fb, x2, dict(opt), None, rng=r2, width=4)
            calls += sp.calls
            assert np.array_equal(x1, x2)
        assert r1.random() == r2.random()
        assert calls < 0.7 * seq_evals[0]            # fewer sequential device calls


def test_slice_sampler_nan_guard(capsys):
    s = samplers.slice()
    out = s(lambda x, a: float("nan") if abs(x[0, 0]) > 0 else 0.0, np.zeros((1, 1)), {"step_out": False}, None,
            rng=np.random.default_rng(1))
    assert "NaN" in capsys.readouterr().out and out.shape == (1, 1)


def test_score_config_defaults():
    ei = scores.expected_improvement()
    assert ei.config == {"tradeoff": 0.0, "nFantasies": 100}
    cb = scores.confidence_bound({"tradeoff": 0})        # 0 is truthy in Lua: stays 0
    assert cb.config["tradeoff"] == 0 and cb.config["bound"] == "lower" and cb.config["sign"] == -1.0
    assert scores.score_args(cb) == (1, 0.0, 0, -1.0)
    assert scores.score_args(scores.confidence_bound({"bound": "upper", "sign": 1.0}))[2:] == (1, 1.0)


def test_c_abi_shard_rule_matches_the_host_rule():
    # b7_shard_range (pure host code in the library: no GPU needed) is the rule that shards candidates AND draws inside
    # b7_gp_fit_sharded / b7_acq_score_multi; the Python twin and bench.py use parallel.shard_range
    import ctypes as C
    from bot7_b200 import _lib as L
    from bot7_b200 import parallel
    lib = L.lib()
    for M in (0, 1, 5, 32, 37888, 2 ** 22 + 3):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for rank in range(world):
                r0, cnt = C.c_int64(), C.c_int64()
                assert lib.b7_shard_range(M, world, rank, C.byref(r0), C.byref(cnt)) == 0
                assert (r0.value, cnt.value) == parallel.shard_range(M, world, rank)
                assert r0.value == covered
                covered += cnt.value
            assert covered == M
    r0, cnt = C.c_int64(), C.c_int64()
    assert lib.b7_shard_range(10, 2, 2, C.byref(r0), C.byref(cnt)) < 0          # rank out of range


def test_comm_init_fails_loudly_without_gpu():
    import ctypes as C
    from bot7_b200 import _lib as L
    if L.lib().b7_device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert L.lib().b7_comm_init_all(2, None, C.byref(h)) < 0
    assert b"devices" in L.lib().b7_last_error() or b"CUDA" in L.lib().b7_last_error()
