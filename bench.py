#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json on the B200-native path.

metric : EI candidates scored per second at N_obs = 4096, S = 32 hyper-parameter draws
         (Hartmann6, d = 6, ARD-SE GP), plus GP fit ms (K build + batched Cholesky) beside it.
step   : one pass of the hot path over one batch of M candidates per GPU: for every draw the K*
         tile, the posterior (V = L^-1 K*^T: exact int8 slice products on tcgen05 by default, FP64 DMMA
         tiles timed beside it), then the fused EI + S-average + argmax pass, and (N > 1) the
         all-gather of the per-rank (best, index, nan) triples.
value  : whole-job candidates/s with the S fitted factors and the candidate grid resident in HBM.
e2e    : the same acquisition through the C ABI with HOST buffers: b7_gp_fit (X, y, hyp from the
         host), b7_grid_from_host (H2D of the batch), b7_acq_score with the score vector read back;
         median of 3 steps after 2 untimed ones.
Launch : python bench.py --gpus N --steps K --warmup W   (torchrun for N > 1, one rank per GPU).
         python bench.py --impl reference ...            (CPU arm: the oracle port on the host cores)
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, DIMS, S_DRAWS = 4096, 6, 32
SMS = 148
PANELS_PER_STEP = 2                       # 2 x (148 SMs x 128 candidates) = 37 888 candidates per GPU per step
M_STEP = PANELS_PER_STEP * SMS * 128
METRIC = "EI candidates scored/s at N=4096, S=32 hypers"
UNIT = "candidates/s"


# ------------------------------------------------------------------ synthetic workload

def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return z ^ (z >> np.uint64(31))


def u01(seed, idx):
    """Counter-based uniform: (splitmix64(seed * 2^32 + i) >> 11) * 2^-53."""
    with np.errstate(over="ignore"):
        i = np.asarray(idx, dtype=np.uint64)
        return (splitmix64(np.uint64(seed) * np.uint64(1 << 32) + i) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def hartmann6(X):
    """benchmarks/hartmann6.lua:36-63 (restated; only synthesises Y_obs)."""
    A = -np.array([[10.0, 3.00, 17.0, 3.50, 1.70, 8.00], [0.05, 10.0, 17.0, 0.10, 8.00, 14.0],
                   [3.00, 3.50, 1.70, 10.0, 17.0, 8.00], [17.0, 8.00, 0.05, 10.0, 0.10, 14.0]])
    P = -np.array([[.1312, .1696, .5569, .0124, .8283, .5886], [.2329, .4135, .8307, .3736, .1004, .9991],
                   [.2348, .1451, .3522, .2883, .3047, .6650], [.4047, .8828, .8732, .5743, .1091, .0381]])
    a = -np.array([1.0, 1.2, 3.0, 3.2])
    return np.exp((A[None] * (X[:, None, :] + P[None]) ** 2).sum(2)) @ a


def hyper_draws(S, d):
    s = np.arange(S)
    hyp = np.zeros((S, d + 3))
    for dd in range(d):
        hyp[:, dd] = np.log(0.1) + u01(2, s * d + dd) * (np.log(2.0) - np.log(0.1))
    hyp[:, d] = 0.5 * (u01(3, s) - 0.5)
    hyp[:, d + 1] = 0.5 * np.log(1e-2)            # noisy case sigma_n^2 = 1e-2
    hyp[:, d + 2] = 0.1 * (u01(4, s) - 0.5)
    return hyp


# ------------------------------------------------------------------ clocks

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop, self.t = index, [], False, None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm (oracle port)

def cpu_scoring_sample(n_draws, m_cand, reps=1):
    """Oracle (numpy/LAPACK/BLAS on all host threads) on a bounded sample of the same workload:
    factors pre-fit, then per draw K* + TRSM + col-sum-sq + EI over m_cand candidates.  Returns
    (seconds per draw for m_cand candidates, fit seconds per draw, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import b7_oracle as o
    try:    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    X = o.sobol_points(DIMS, N_OBS + m_cand)
    Xo, Xc = X[:N_OBS], X[N_OBS:]
    y = o.hartmann6(Xo)
    y = (y - y.mean()) / y.std()
    hyp = hyper_draws(S_DRAWS, DIMS)[:n_draws]
    t0 = time.perf_counter()
    fits = [o.gp_fit(Xo, y, h, 0, False) for h in hyp]
    t_fit = (time.perf_counter() - t0) / n_draws
    fmin = float(y.min())
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        acc = np.zeros(m_cand)
        for f in fits:
            mu, var = o.gp_predict(f, Xc)
            acc = acc + o.ei_compute(mu, var, fmin, 0.0)
        o.argmax_first(acc / n_draws)
        best = min(best, (time.perf_counter() - t0) / n_draws)
    return best, t_fit, os.cpu_count()


def run_reference(args, rank):
    if rank != 0:
        return
    n_draws, m_cand = 2, 2048
    times = []
    t_fit = None
    for i in range(args.warmup + args.steps):
        t, t_fit, cores = cpu_scoring_sample(n_draws, m_cand)
        if i >= args.warmup:
            times.append(t)
    per_draw = statistics.mean(times)
    value = m_cand / (per_draw * S_DRAWS)          # work is exactly linear in S
    sample = (f"{n_draws} of {S_DRAWS} draws x {m_cand} candidates per step, factors pre-fit; candidates/s scaled linearly to "
              f"S={S_DRAWS}; numpy/scipy (OpenBLAS) on all host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_draw * n_draws * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Hartmann6 integrated EI, N_obs=4096, d=6, S=32 draws (headline of BASELINE.json metric)",
                       "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "fit_ms_per_factor": t_fit * 1e3,
                             "note": "CPU restatement (Torch7/gpTorch7 unavailable: no Lua runtime, gp rock not vendored)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.warmup < 3:
        args.warmup = 3

    from bot7_b200 import _lib as L
    from bot7_b200 import grids, models, parallel
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = L.Context.default(local)
    lib = L.lib()

    # observations: first N_OBS Sobol points; candidates: this rank's shard of the rest
    sob = grids.sobol({"size": N_OBS + world * M_STEP, "dims": DIMS})
    Xo = sob.generate({"size": N_OBS, "dims": DIMS})
    y = hartmann6(Xo)
    y = (y - y.mean()) / y.std()
    hyp = hyper_draws(S_DRAWS, DIMS)
    fmin = float(y.min())
    row0, cnt = parallel.shard_range(world * M_STEP, world, rank)
    grid = sob.generate_device(first=N_OBS + row0, count=cnt)

    # ---- fit (timed separately: "GP fit ms") -----------------------------------------------------
    ctx.set_profiling(True)
    fit_ms = {}
    for rep in range(2):
        ctx.reset_timers()
        if world > 1 and S_DRAWS >= world:
            f = models.GPFactors(Xo, y, hyp, "ardse", False, L.FIT_DEFER, ctx)
            s0, sc = parallel.draw_range(S_DRAWS, world, rank)
            f.fit_range(s0, sc)
            t0 = time.perf_counter()
            parallel.allgather_factors(f, S_DRAWS, world, rank)
            fit_ms["allgather_ms"] = (time.perf_counter() - t0) * 1e3
            f.mark_ready()
        else:
            f = models.GPFactors(Xo, y, hyp, "ardse", False, L.FIT_PREDICT, ctx)
        st = ctx.stage_times()
        fit_ms.update(kbuild_ms=st["kbuild"][0], potrf_ms=st["potrf"][0], trtri_ms=st["trtri"][0],
                      draws_on_this_rank=(S_DRAWS if world == 1 or S_DRAWS < world else parallel.draw_range(S_DRAWS, world, rank)[1]))
        if rep == 0:
            f.free()
    assert (f.info == 0).all()
    # single-factor density evaluation (the slice sampler's f): X, y resident, new hyper-parameters each call
    f1 = models.GPFactors(Xo, y, hyp[:1], "ardse", False, L.FIT_LOGML_ONLY, ctx)
    for i in range(2):
        f1.refit(hyp[1:2], L.FIT_LOGML_ONLY)
    t0 = time.perf_counter()
    for i in range(5):
        f1.refit(hyp[i:i + 1], L.FIT_LOGML_ONLY)
    fit_ms["single_factor_refit_ms"] = (time.perf_counter() - t0) / 5 * 1e3
    f1.free()

    def step():
        o_, b_, n_ = C.c_int64(), C.c_double(), C.c_int64()
        L.check(lib.b7_acq_score_range(f.handle, grid.handle, 0, cnt, L.SCORE_EI, 0.0, 0, -1.0, fmin, None, C.byref(o_),
                                       C.byref(b_), C.byref(n_)), "b7_acq_score_range")
        gi = o_.value + row0 if o_.value > 0 else 0
        return parallel.allgather_argmax(b_.value, gi, n_.value) if world > 1 else (b_.value, gi, n_.value)

    def timed(path, warmup, steps, profiled):
        """`steps` timed steps on the given posterior path; returns (device ms per step (max over ranks), wall ms, launches, stage times,
        clocks, result).  profiled = False: the number that counts (no per-stage synchronisation, K* of draw s + 1 overlaps the posterior
        pass of draw s); profiled = True: per-launch CUDA-event times for the roofline (every stage synchronises)."""
        ctx.set_posterior_path(path)
        ctx.set_profiling(profiled)
        for _ in range(warmup):
            r = step()
        if dist:
            dist.barrier()
        ctx.sync()
        ctx.reset_timers()
        l0 = ctx.launch_count()
        with ClockSampler(local) as clk_:
            ctx.timer_begin()
            t0 = time.perf_counter()
            for _ in range(steps):
                r = step()
            dev = ctx.timer_end()
            wall = (time.perf_counter() - t0) * 1e3
        if dist:
            dist.barrier()
        n_l = ctx.launch_count() - l0
        st_ = ctx.stage_times()
        ctx.set_profiling(False)
        if dist:
            import torch
            t = torch.tensor([dev, wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev, wall = float(t[0]), float(t[1])
        return dev / steps, wall / steps, n_l, st_, clk_.summary(), r

    default_path = ctx.posterior_path()
    other_path = L.PATH_FP64_DMMA if default_path == L.PATH_INT8_OZAKI else L.PATH_INT8_OZAKI
    # secondary path first (so that the default path is the one left selected), then the default path = `value`
    o_ms, o_wall, o_launches, _, o_clk, o_res = timed(other_path, 2, args.steps, False)
    _, _, _, o_st, _, _ = timed(other_path, 0, 1, True)
    ms_per_step, wall_ms, launches, _, clocks, res = timed(default_path, args.warmup, args.steps, False)
    prof_ms, _, _, st, prof_clk, _ = timed(default_path, 0, 2, True)
    PROF_STEPS, O_PROF_STEPS = 2, 1
    value = world * M_STEP / (ms_per_step * 1e-3)
    other_value = world * M_STEP / (o_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers (fit + upload + score + read back) ------
    ctx.set_profiling(False)
    Xc_host = grid.read()
    e2e_times = []
    # two untimed iterations (the stream-ordered pool still grows: fit 920 -> 104 -> 35 ms, tools/e2e_breakdown.py), then three
    for i in range(5):
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        f2 = models.GPFactors(Xo, y, hyp, "ardse", False, L.FIT_PREDICT, ctx)
        g2 = grids.DeviceGrid.from_host(Xc_host, ctx)
        score = np.empty(cnt)
        am, amo, b_, n_ = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
        L.check(lib.b7_acq_score(f2.handle, g2.handle, L.SCORE_EI, 0.0, 0, -1.0, fmin, L.dptr(score), C.byref(am), C.byref(amo),
                                 C.byref(b_), C.byref(n_)), "b7_acq_score")
        if world > 1:
            parallel.allgather_argmax(b_.value, amo.value + row0, n_.value)
        e2e_times.append(time.perf_counter() - t0)
        f2.free()
        g2.free()
    e2e_s = float(np.median(e2e_times[2:]))
    if dist:
        import torch
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    h2d = Xc_host.nbytes + Xo.nbytes + y.nbytes + hyp.nbytes
    d2h = cnt * 8 + 32

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (posterior TRMM, FP64 tensor pipe) ----------------------
    peak_tf, peak_src = 37.1, "fallback constant"
    try:
        pk = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))
        peak_tf = max(v for k, v in pk.items() if k.startswith("dmma_"))
        peak_src = "profiles/fp64_peak_r01.json: DMMA.8x8x4 issue-rate probe measured on this pool's B200 (tools/fp64_peak.cu); " \
                   "MEASURED_PEAKS.json carries no FP64 figure"
    except Exception:
        pass
    hbm = 6451.8
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    i8_peak, i8_src = 4010.5, "fallback constant"
    try:
        pk8 = json.load(open(os.path.join(ROOT, "profiles", "i8_mma_peak_r01.json")))
        i8_peak = pk8["i8_mma_m128n128_tops"]
        i8_src = "profiles/i8_mma_peak_r01.json: tcgen05.mma.kind::i8 M128 N128 issue-rate probe on this pool's B200 (tools/i8_mma_probe.cu)"
    except Exception:
        pass
    flops_per_launch = (SMS * 128) * (float(N_OBS) ** 2 + 4.0 * N_OBS)      # one panel x one draw

    def posterior_roofline(path, st_, ms_step, n_steps):
        post_ms, post_n = st_["posterior"]
        if not post_n:
            return None
        avg = post_ms / post_n
        if path == L.PATH_FP64_DMMA:
            ach = flops_per_launch / (avg * 1e-3) * 1e-12
            return {"bound": "tensor", "kernel": "posterior_kernel (FP64 DMMA.8x8x4)", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf, "traffic": traffic, "peak_source": peak_src, "launches_timed": post_n, "avg_launch_ms": avg,
                    "share_of_step": post_ms / n_steps / ms_step}
        ach_eq = flops_per_launch / (avg * 1e-3) * 1e-12
        ach_i8 = 28.0 * ach_eq                                                # 28 exact int8 slice products per fp64 product
        return {"bound": "tensor", "kernel": "posterior_i8_kernel (tcgen05.mma.kind::i8, 7x7 error-free radix-256 slices, 28 products)",
                "achieved": ach_i8, "peak": i8_peak, "unit": "TOP/s", "frac": ach_i8 / i8_peak, "traffic": traffic_i8,
                "fp64_equivalent_tflops": ach_eq, "fp64_dmma_peak_tflops": peak_tf, "peak_source": i8_src,
                "note": "N = 64 MMAs (7 int32 accumulators x 64 columns = 448 of the 512 TMEM columns) with the A operand held in the "
                        "collector reach 3814 TOP/s in the same probe (2754 without the collector: shared-memory operand reads); "
                        "peak is the N = 128 figure",
                "launches_timed": post_n, "avg_launch_ms": avg, "share_of_step": post_ms / n_steps / ms_step,
                "timing": "per-launch CUDA events in a separate profiled pass (every stage synchronised); share_of_step = posterior ms per "
                          "profiled step / ms_per_step of the unprofiled timed loop"}

    traffic = traffic_i8 = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "posterior_ncu_r01.json"))).get("dram_bytes_per_launch")
        traffic_i8 = json.load(open(os.path.join(ROOT, "profiles", "posterior_i8_ncu_r01.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = posterior_roofline(default_path, st, ms_per_step, PROF_STEPS)
    roofline_other = posterior_roofline(other_path, o_st, o_ms, O_PROF_STEPS)
    path_name = {L.PATH_FP64_DMMA: "fp64_dmma", L.PATH_INT8_OZAKI: "int8_ozaki"}
    Np = N_OBS
    stages = {
        "profiled_ms_per_step": prof_ms,
        "kstar": {"ms_per_step": st["kstar"][0] / PROF_STEPS, "bound": "hbm",
                  "achieved_gbs": (M_STEP * S_DRAWS * Np * 8) / (st["kstar"][0] / PROF_STEPS * 1e-3) * 1e-9, "peak_gbs": hbm,
                  "note": "serialised in the profiled pass; in the timed loop it runs on a second stream under the posterior pass"},
        "score": {"ms_per_step": st["score"][0] / PROF_STEPS, "bound": "hbm",
                  "achieved_gbs": (M_STEP * (16 * S_DRAWS + 8)) / (st["score"][0] / PROF_STEPS * 1e-3) * 1e-9, "peak_gbs": hbm},
        "fit_kbuild": {"ms": fit_ms["kbuild_ms"], "bound": "hbm",
                       "achieved_gbs": fit_ms["draws_on_this_rank"] * Np * Np * 8 / (fit_ms["kbuild_ms"] * 1e-3) * 1e-9, "peak_gbs": hbm},
        # fp64-equivalent rates: on the int8_ozaki path the k = 512 trailing updates (potrf_i8.cu) and the inversion
        # (trtri_i8.cu) run as exact int8 slice products, so the figure can exceed the FP64 DMMA peak it is shown next to
        "fit_potrf": {"ms": fit_ms["potrf_ms"], "bound": "tensor",
                      "achieved_tflops": fit_ms["draws_on_this_rank"] * Np ** 3 / 3 / (fit_ms["potrf_ms"] * 1e-3) * 1e-12, "peak_tflops": peak_tf,
                      "arithmetic": "fp64 DMMA + int8 slices (k=512 updates)" if default_path == L.PATH_INT8_OZAKI else "fp64 DMMA"},
        "fit_trtri": {"ms": fit_ms["trtri_ms"], "bound": "tensor",
                      "achieved_tflops": fit_ms["draws_on_this_rank"] * Np ** 3 / 3 / (fit_ms["trtri_ms"] * 1e-3) * 1e-12, "peak_tflops": peak_tf,
                      "arithmetic": "int8 slices, block-recursive" if default_path == L.PATH_INT8_OZAKI else "fp64 DMMA"},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "Hartmann6 integrated EI, N_obs=4096, d=6, S=32 draws (headline of BASELINE.json metric; fits one GPU)",
                   "arithmetic": "fp64 results; on the int8_ozaki path the N^2 products are 28 exact int8 slice products recombined in fp64",
                   "kernel": "ARD-SE", "candidates_per_gpu_per_step": M_STEP, "grid": "Sobol (generated on device, per-rank shard)",
                   "l2": "inputs larger than L2: 32 inverse factors = 4.3 GB + 620 MB K* panel per launch vs 126 MB L2",
                   "parallelism": f"candidate-sharded x{world}" + (", draw-sharded fit + NCCL all-gather" if world > 1 else "")},
        "gp_fit_ms": {"k_build_plus_cholesky_S32": fit_ms["kbuild_ms"] + fit_ms["potrf_ms"],
                      "per_factor": (fit_ms["kbuild_ms"] + fit_ms["potrf_ms"]) / fit_ms["draws_on_this_rank"],
                      "inversion_for_predict_S32": fit_ms["trtri_ms"], **{k: v for k, v in fit_ms.items() if k.endswith("_ms")}},
        "e2e": {"value": world * cnt / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "includes": "b7_gp_fit from host X/y/hyp (K build, batched potrf, inversion) + b7_grid_from_host + b7_acq_score "
                            "with the score vector copied back; median of 3 steps after 2 untimed ones", "seconds_per_step": e2e_s,
                "all_steps_s": [round(x, 4) for x in e2e_times]},
        "gpu_launches": int(launches), "wall_ms_per_step": wall_ms,
        "clocks": clocks, "roofline": roofline, "stages": stages,
        "posterior_path": path_name[default_path],
        "other_path": {"posterior_path": path_name[other_path], "value": other_value, "unit": UNIT, "ms_per_step": o_ms,
                       "gpu_launches": int(o_launches), "roofline": roofline_other, "clocks": o_clk,
                       "same_argmax": bool(o_res[1] == res[1]), "best_rel_diff": abs(o_res[0] - res[0]) / abs(res[0]) if res[0] else None},
        "result": {"best": res[0], "global_index": res[1], "nan_count": res[2]},
    }
    if not args.no_cpu_baseline:
        t_draw, t_fit, cores = cpu_scoring_sample(2, 2048)
        line["cpu_baseline"] = {"value": 2048 / (t_draw * S_DRAWS), "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "2 of 32 draws x 2048 candidates (factors pre-fit), scaled linearly to S=32; "
                                          "numpy/scipy OpenBLAS on all host threads",
                                "fit_ms_per_factor": t_fit * 1e3,
                                "note": "CPU restatement (Torch7/gpTorch7 unavailable)"}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    if not line["other_path"]["same_argmax"]:
        print("bench: the two posterior paths selected different candidates", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
