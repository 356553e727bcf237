"""An oracle-backed stand-in for libbot7_b200.so -- TEST INFRASTRUCTURE for the CPU-side protocol test only.

tests/test_lua_reference_loop.py runs the reference's own control plane (init.lua, bots/abstract.lua, bots/bayesopt.lua,
samplers/slice.lua, examples/run_benchmark.lua ...) with the LuaJIT glue installed, under tools/minilua, in the build container --
which has the reference tree but no GPU.  There the glue's `ffi.load` is answered by this object: every C entry point the GP
path of the glue calls, implemented with the CPU oracle on the memory the glue passes (same signatures as include/bot7_b200.h:
the FFI stand-in converts the arguments exactly as for the real library and hands pointers over as addresses).  It exists to
check the PROTOCOL between the reference's classes, the glue and the C ABI (who owns which field, which index numbering is used
when, what gets freed) -- not the numerics, which the GPU tests check against the same oracle through the real library.
It is never importable from the product: it lives under tests/ and needs oracle/.
"""
import ctypes as C

import numpy as np


def _arr(addr, shape, ctype=C.c_double):
    n = int(np.prod(shape))
    return np.ctypeslib.as_array((ctype * n).from_address(addr)).reshape(shape)


class FakeB7:
    def __init__(self, oracle):
        self.o = oracle
        self.handles = {}
        self.next_id = 0x1000
        self.calls = []
        self.freed = []
        self.err = b""
        self._keep = []
        self._fns = {}
        for name in dir(type(self)):
            if name.startswith("b7_"):
                # a plain function in the instance dictionary (shadows the method): the FFI stand-in sets restype / argtypes on it
                self.__dict__[name] = self._fns[name] = self._wrap(name, getattr(self, name))

    def _wrap(self, name, method):
        def fn(*a):
            self.calls.append(name)
            try:
                return method(*a)
            except FakeError as e:
                self.err = str(e).encode()
                return -1
        fn.__name__ = name
        return fn

    def __getattr__(self, name):
        fns = self.__dict__.get("_fns", {})
        if name in fns:
            return fns[name]
        raise AttributeError(name)

    def _new(self, obj):
        self.next_id += 0x10
        self.handles[self.next_id] = obj
        return self.next_id

    def _get(self, h, kind):
        obj = self.handles.get(h)
        if obj is None or obj["kind"] != kind:
            raise FakeError("%s handle expected (got %r: %s)" % (kind, h, "freed or foreign" if obj is None else obj["kind"]))
        return obj

    # ---- context
    def b7_version(self):
        return 100

    def b7_last_error(self):
        return self.err                               # (the FFI stand-in wraps a const char* result as bytes)

    def b7_device_count(self):
        return 1

    def b7_init(self, device, out):
        C.c_void_p.from_address(out).value = self._new({"kind": "ctx", "device": device})
        return 0

    def b7_shutdown(self, ctx):
        live = [h for h, o in self.handles.items() if o["kind"] != "ctx"]
        self.freed.append(("ctx", ctx, len(live)))
        self.handles.pop(ctx, None)

    # ---- grids
    def b7_sobol_generate(self, ctx, dims, first_seed, count, mins, maxes, out_host, out_grid):
        self._get(ctx, "ctx")
        lo = _arr(mins, (dims,)).copy() if mins else None
        hi = _arr(maxes, (dims,)).copy() if maxes else None
        pts = self.o.sobol_points(dims, count, first_seed, lo, hi) if (lo is not None and hi is not None) else self.o.sobol_points(dims, count, first_seed)
        if (lo is None) != (hi is None):
            raise FakeError("sobol_generate: mins and maxes come together (the one-sided variants are done by the caller)")
        if out_host:
            _arr(out_host, (count, dims))[...] = pts
        if out_grid:
            C.c_void_p.from_address(out_grid).value = self._new({"kind": "grid", "X": pts.copy(), "live": list(range(count))})
        return 0

    def b7_grid_from_host(self, ctx, X, M, d, out_grid):
        self._get(ctx, "ctx")
        C.c_void_p.from_address(out_grid).value = self._new({"kind": "grid", "X": _arr(X, (M, d)).copy(), "live": list(range(M))})
        return 0

    def b7_grid_size(self, g):
        return len(self._get(g, "grid")["live"])

    def b7_grid_remove(self, g, idx, removed_row):
        grid = self._get(g, "grid")
        if not 1 <= idx <= len(grid["live"]):
            raise FakeError("grid_remove: index %d out of range (live rows %d)" % (idx, len(grid["live"])))
        orig = grid["live"].pop(idx - 1)
        if removed_row:
            _arr(removed_row, (grid["X"].shape[1],))[...] = grid["X"][orig]
        return 0

    def b7_grid_free(self, g):
        self._get(g, "grid")
        self.freed.append(("grid", g))
        del self.handles[g]

    # ---- GP
    def _fit(self, gp, hyp, info, logml, jitter):
        fits = []
        for s in range(gp["S"]):
            f = self.o.gp_fit(gp["X"], gp["y"], hyp[s], gp["kernel"], gp["noiseless"])
            fits.append(f)
            if info:
                _arr(info, (gp["S"],), C.c_int)[s] = 0
            if logml:
                _arr(logml, (gp["S"],))[s] = f["logml"]
            if jitter:
                _arr(jitter, (gp["S"],))[s] = f.get("jitter", 0.0)
        gp["fits"], gp["hyp"] = fits, hyp.copy()

    def b7_gp_fit(self, ctx, kernel, X, y, N, d, hyp, S, H, noiseless, flags, out, info, logml, jitter):
        self._get(ctx, "ctx")
        if H != d + 3:
            raise FakeError("gp_fit: H = %d, expected d + 3 = %d" % (H, d + 3))
        gp = {"kind": "gp", "X": _arr(X, (N, d)).copy(), "y": _arr(y, (N,)).copy(), "kernel": kernel, "noiseless": bool(noiseless),
              "S": S, "flags": flags}
        self._fit(gp, _arr(hyp, (S, H)).copy(), info, logml, jitter)
        C.c_void_p.from_address(out).value = self._new(gp)
        return 0

    def b7_gp_refit(self, gp, hyp, flags, info, logml, jitter):
        g = self._get(gp, "gp")
        g["flags"] = flags
        self._fit(g, _arr(hyp, (g["S"], g["X"].shape[1] + 3)).copy(), info, logml, jitter)
        return 0

    def b7_gp_predict(self, gp, s, Xs, M, mean, var):
        g = self._get(gp, "gp")
        if g["flags"] == 1:
            raise FakeError("gp_predict: the handle was fitted with B7_FIT_LOGML_ONLY")
        mu, v = self.o.gp_predict(g["fits"][s], _arr(Xs, (M, g["X"].shape[1])).copy())
        _arr(mean, (M,))[...] = mu
        _arr(var, (M,))[...] = v
        return 0

    def b7_gp_free(self, gp):
        self._get(gp, "gp")
        self.freed.append(("gp", gp))
        del self.handles[gp]

    # ---- acquisition
    def b7_acq_score(self, gp, grid, kind, tradeoff, bound, sign, fmin, score_host, argmax, argmax_original, best, nan_count):
        g, gr = self._get(gp, "gp"), self._get(grid, "grid")
        if g["flags"] == 1:
            raise FakeError("acq_score: the handle was fitted with B7_FIT_LOGML_ONLY")
        Xs = gr["X"][gr["live"]]
        per = []
        for f in g["fits"]:
            mu, v = self.o.gp_predict(f, Xs)
            per.append(self.o.ei_compute(mu, v, fmin, tradeoff) if kind == 0 else
                       self.o.cb_compute(mu, v, tradeoff, "upper" if bound == 1 else "lower", sign))
        score = self.o.mc_average(np.array(per))
        b, idx, nans = self.o.argmax_first(score)
        if score_host:
            _arr(score_host, (len(score),))[...] = score
        C.c_int64.from_address(argmax).value = idx
        if argmax_original:
            C.c_int64.from_address(argmax_original).value = gr["live"][idx - 1] + 1 if idx else 0
        C.c_double.from_address(best).value = b
        C.c_int64.from_address(nan_count).value = nans
        return 0

    def b7_score_moments(self, ctx, kind, mean, var, S, M, tradeoff, bound, sign, fmin, score_host, argmax, best, nan_count):
        self._get(ctx, "ctx")
        m, v = _arr(mean, (S, M)), _arr(var, (S, M))
        per = [self.o.ei_compute(m[s], v[s], fmin, tradeoff) if kind == 0 else
               self.o.cb_compute(m[s], v[s], tradeoff, "upper" if bound == 1 else "lower", sign) for s in range(S)]
        score = self.o.mc_average(np.array(per))
        _arr(score_host, (M,))[...] = score
        b, idx, nans = self.o.argmax_first(score)
        if argmax:
            C.c_int64.from_address(argmax).value = idx
        if best:
            C.c_double.from_address(best).value = b
        if nan_count:
            C.c_int64.from_address(nan_count).value = nans
        return 0


    # ---- DNGO: basis + Bayesian linear regression head
    def _stack(self, n_layers, dims, W, b):
        d = _arr(dims, (n_layers + 1,), C.c_int)
        Ws, bs = [], []
        for l in range(n_layers):
            wp = C.c_void_p.from_address(W + l * C.sizeof(C.c_void_p)).value
            bp = C.c_void_p.from_address(b + l * C.sizeof(C.c_void_p)).value
            Ws.append(_arr(wp, (int(d[l + 1]), int(d[l]))).copy())
            bs.append(_arr(bp, (int(d[l + 1]),)).copy())
        return Ws, bs

    def b7_grid_read(self, g, first, count, out_host):
        grid = self._get(g, "grid")
        _arr(out_host, (count, grid["X"].shape[1]))[...] = grid["X"][first:first + count]
        return 0

    def b7_mlp_features(self, ctx, g, n_layers, dims, W, b, relu_last, out):
        self._get(ctx, "ctx")
        grid = self._get(g, "grid")
        Ws, bs = self._stack(n_layers, dims, W, b)
        Z = self.o.mlp_features(grid["X"], Ws, bs, bool(relu_last))
        C.c_void_p.from_address(out).value = self._new({"kind": "grid", "X": Z, "live": list(grid["live"])})
        return 0

    def b7_blr_fit(self, ctx, Z0, y, N, D, hyp, S, out, info):
        self._get(ctx, "ctx")
        Z, yy, H = _arr(Z0, (N, D)).copy(), _arr(y, (N,)).copy(), _arr(hyp, (S, 3)).copy()
        fits = [self.o.blr_fit(Z, yy, H[s]) for s in range(S)]
        if info:
            _arr(info, (S,), C.c_int)[...] = 0
        C.c_void_p.from_address(out).value = self._new({"kind": "blr", "fits": fits, "D": D})
        return 0

    def b7_blr_predict(self, blr, s, Z1, M, mean, var):
        f = self._get(blr, "blr")
        mu, v = self.o.blr_predict(f["fits"][s], _arr(Z1, (M, f["D"])).copy())
        _arr(mean, (M,))[...] = mu
        _arr(var, (M,))[...] = v
        return 0

    def b7_dngo_score(self, blr, g, n_layers, dims, W, b, relu_last, kind, tradeoff, bound, sign, fmin, score_host, argmax,
                      argmax_original, best, nan_count):
        f, grid = self._get(blr, "blr"), self._get(g, "grid")
        Ws, bs = self._stack(n_layers, dims, W, b)
        Z1 = self.o.mlp_features(grid["X"][grid["live"]], Ws, bs, bool(relu_last))
        per = []
        for fit in f["fits"]:
            mu, v = self.o.blr_predict(fit, Z1)
            per.append(self.o.ei_compute(mu, v, fmin, tradeoff) if kind == 0 else
                       self.o.cb_compute(mu, v, tradeoff, "upper" if bound == 1 else "lower", sign))
        score = self.o.mc_average(np.array(per))
        bst, idx, nans = self.o.argmax_first(score)
        if score_host:
            _arr(score_host, (len(score),))[...] = score
        C.c_int64.from_address(argmax).value = idx
        if argmax_original:
            C.c_int64.from_address(argmax_original).value = grid["live"][idx - 1] + 1 if idx else 0
        C.c_double.from_address(best).value = bst
        C.c_int64.from_address(nan_count).value = nans
        return 0

    def b7_blr_free(self, blr):
        self._get(blr, "blr")
        self.freed.append(("blr", blr))
        del self.handles[blr]


    # ---- multi-GPU block: one process, n devices; shards are index ranges of ONE logical grid with a shared tombstone list
    def _shards(self, M, n):
        base, rem = divmod(M, n)
        out, r0 = [], 0
        for g in range(n):
            cnt = base + (1 if g < rem else 0)
            out.append((r0, cnt))
            r0 += cnt
        return out

    def b7_comm_init_all(self, n_gpus, device_ids, out):
        if n_gpus < 1:
            raise FakeError("comm_init_all: n_gpus")
        C.c_void_p.from_address(out).value = self._new({"kind": "comm", "n": n_gpus})
        return 0

    def b7_comm_free(self, comm):
        self._get(comm, "comm")
        self.freed.append(("comm", comm))
        del self.handles[comm]

    def b7_grid_from_host_sharded(self, comm, X, M, d, out_grids):
        c = self._get(comm, "comm")
        whole = {"X": _arr(X, (M, d)).copy(), "live": list(range(M))}          # shared by the shard handles
        for g, (r0, cnt) in enumerate(self._shards(M, c["n"])):
            h = self._new({"kind": "grid", "X": whole["X"], "live": whole["live"], "whole": whole, "shard": (r0, cnt)})
            C.c_void_p.from_address(out_grids + g * C.sizeof(C.c_void_p)).value = h
        return 0

    def b7_grid_remove_sharded(self, comm, grids, idx, removed_row):
        c = self._get(comm, "comm")
        hs = [C.c_void_p.from_address(grids + g * C.sizeof(C.c_void_p)).value for g in range(c["n"])]
        first = self._get(hs[0], "grid")
        for h in hs[1:]:
            if self._get(h, "grid")["whole"] is not first["whole"]:
                raise FakeError("grid_remove_sharded: the handles do not belong to one sharded grid")
        return FakeB7.b7_grid_remove(self, hs[0], idx, removed_row)

    def b7_gp_fit_sharded(self, comm, kernel, X, y, N, d, hyp, S, H, noiseless, out_gps, info, logml, jitter, gather_ms):
        c = self._get(comm, "comm")
        gp = {"kind": "gp", "X": _arr(X, (N, d)).copy(), "y": _arr(y, (N,)).copy(), "kernel": kernel, "noiseless": bool(noiseless),
              "S": S, "flags": 0}
        self._fit(gp, _arr(hyp, (S, H)).copy(), info, logml, jitter)
        for g in range(c["n"]):                                                # every device ends up with all S factors
            C.c_void_p.from_address(out_gps + g * C.sizeof(C.c_void_p)).value = self._new(dict(gp))
        if gather_ms:
            C.c_double.from_address(gather_ms).value = 0.0
        return 0

    def b7_acq_score_multi(self, comm, gps, grids, kind, tradeoff, bound, sign, fmin, score_host, argmax, argmax_original, best, nan_count):
        c = self._get(comm, "comm")
        gh = [C.c_void_p.from_address(gps + g * C.sizeof(C.c_void_p)).value for g in range(c["n"])]
        rh = [C.c_void_p.from_address(grids + g * C.sizeof(C.c_void_p)).value for g in range(c["n"])]
        for h in gh:
            self._get(h, "gp")
        for h in rh:
            self._get(h, "grid")
        return FakeB7.b7_acq_score(self, gh[0], rh[0], kind, tradeoff, bound, sign, fmin, score_host, argmax, argmax_original, best, nan_count)


class FakeError(Exception):
    pass
