"""bot7.bots -- experiment loop (host-side mirror of reference bots/abstract.lua, bots/bayesopt.lua,
bots/random_search.lua).  The control flow, defaults and bookkeeping follow the reference; the
hot body of bot:eval / bot:nominate (bots/bayesopt.lua:56-99) is one batched device call:
S hyper draws -> S factors (b7_gp_fit) -> fused posterior + score + average + argmax over the
device-resident grid (b7_acq_score).
"""
from __future__ import annotations

import copy
import ctypes as C

import numpy as np

from . import _lib as L
from . import grids as Grids
from . import models as Models
from . import parallel
from . import scores as Scores
from . import t7


class abstract:
    def __init__(self, objective, hypers, config=None, cache=None, ctx=None, rng=None):
        cache = cache or {}
        self.ctx = ctx
        self.rng = rng or np.random.default_rng(0)
        self.hypers = cache.get("hypers", hypers)
        self.objective = cache.get("objective", objective)
        self.config = self.configure(cache.get("config", config))
        config = self.config
        # candidate grid (bots/abstract.lua:31-36): device resident; `candidates` materialises it
        self.grid = cache.get("grid")
        if self.grid is None:
            if cache.get("candidates") is not None:
                self.grid = Grids.DeviceGrid.from_host(cache["candidates"], ctx)
            else:
                gtype = config["grid"]["type"]
                if gtype == "random":
                    self.grid = Grids.random(config["grid"], self.rng).generate_device(ctx=ctx)
                else:
                    self.grid = getattr(Grids, gtype)(config["grid"], ctx).generate_device()
        self.responses = cache.get("responses")
        self.observed = cache.get("observed")
        self.pending = None
        self.removed_original = []     # 1-based original rows already nominated
        self.nTrials = 0 if self.observed is None else self.observed.shape[0]
        self.best = {"x": np.empty((1, config["grid"]["dims"])), "t": -1}
        if self.responses is not None:
            self.best["y"] = self.responses.min(0)
            self.best["t"] = int(self.responses[:, 0].argmin()) + 1

    @property
    def candidates(self):
        """The compacted candidate tensor of the reference (host copy; only built on request)."""
        full = self.grid.read()
        keep = np.ones(full.shape[0], dtype=bool)
        keep[[r - 1 for r in self.removed_original]] = False
        return full[keep]

    def configure(self, config):
        """bots/abstract.lua:56-109."""
        config = copy.deepcopy(config) if config else {}
        bot = config.get("bot") or {}
        for k, v in (("verbose", 3), ("budget", 100), ("msg_freq", 1), ("nInitial", 2), ("nSamples", 10), ("save", False)):
            if bot.get(k) is None:
                bot[k] = v
        config["bot"] = bot
        score = config.get("score") or {}
        score.setdefault("type", "expected_improvement")
        config["score"] = score
        grid = config.get("grid") or {}
        grid.setdefault("type", "sobol")
        grid.setdefault("size", int(2e4))
        if grid.get("dims") is None:
            grid["dims"] = int(sum(h["size"] for h in self.hypers))
        if grid.get("mins") is None:
            grid["mins"] = np.concatenate([np.broadcast_to(np.asarray(h["min"], float), (h["size"],)) for h in self.hypers])
        if grid.get("maxes") is None:
            grid["maxes"] = np.concatenate([np.broadcast_to(np.asarray(h["max"], float), (h["size"],)) for h in self.hypers])
        config["grid"] = grid
        return config

    def run_trial(self):
        """bots/abstract.lua:112-152."""
        self.nTrials += 1
        idx = int(self.nominate())
        self.removed_original.append(self.grid.original_index(idx))
        nominee = self.grid.remove(idx)                    # steal(pending, candidates, idx)
        self.pending = nominee
        y = np.asarray(self.objective(nominee[0]), dtype=np.float64).reshape(1, -1)
        self.responses = y if self.responses is None else np.concatenate([self.responses, y], 0)
        self.observed = nominee if self.observed is None else np.concatenate([self.observed, nominee], 0)
        self.pending = None
        if getattr(self, "model", None) is not None and self.nTrials == self.config["bot"]["nInitial"]:
            self.model.init(self.observed, self.responses)
        return nominee, y

    def update_best(self, x, y):
        if "y" not in self.best or (self.best["y"] > y).all():
            self.best.update(t=self.nTrials, x=x, y=y)

    def run_experiment(self):
        for t in range(1, self.config["bot"]["budget"] + 1):
            x, y = self.run_trial()
            self.update_best(x, y)
            self.progress_report(t, x, y)
        return self.best

    def progress_report(self, t, x, y):
        c = self.config["bot"]
        if c["verbose"] < 1 or t % c["msg_freq"] != 0:
            return
        print("Trial: %d of %d" % (t, c["budget"]))
        print("> Best response (#%d): %s" % (self.best["t"], np.array2string(np.asarray(self.best["y"]).ravel())))
        print("> Last response (#%d): %s" % (t, np.array2string(np.asarray(y).ravel())))

    def eval(self):
        print("Error: eval() method not implemented")

    def nominate(self):
        print("Error: nominate() method not implemented")

    def class_(self):
        """bots/abstract.lua:246-252: torch.type(self)."""
        return "bot7.bots." + self.__class__.__name__

    def save(self, path=None):
        """bots/abstract.lua:234-240: torch.save('demo_<class>.t7', {best=, x=observed, y=responses}), same file format."""
        best = {k: (np.asarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v) for k, v in self.best.items()}
        results = {"best": best, "x": self.observed, "y": self.responses}
        path = path or "demo_%s.t7" % self.class_()
        t7.save(path, results)
        return path

    __call__ = run_experiment


def cache_from_results(path_or_table, candidates=None):
    """Turns a result file written by bot:save (here or by the reference) into the `cache` argument of a bot
    constructor (bots/abstract.lua:19-44): the run resumes with its observations; `candidates`, if given, is
    the grid to continue on (the reference keeps it in memory only)."""
    res = t7.load(path_or_table) if isinstance(path_or_table, (str, bytes)) or hasattr(path_or_table, "__fspath__") else path_or_table
    x, y = res.get("x"), res.get("y")
    if x is None or y is None:
        raise ValueError("result table has no observations (x, y)")
    x = np.array(x, dtype=np.float64, ndmin=2)
    y = np.array(y, dtype=np.float64)
    if y.ndim == 1:
        y = y.reshape(-1, 1)
    if x.shape[0] != y.shape[0]:
        raise ValueError("result table: %d observed points but %d responses" % (x.shape[0], y.shape[0]))
    cache = {"observed": x, "responses": y}
    if candidates is not None:
        cache["candidates"] = candidates
    return cache


class random_search(abstract):
    def nominate(self):
        """bots/random_search.lua:28-30."""
        return int(self.rng.random() * self.grid.size()) + 1


class bayesopt(abstract):
    def __init__(self, objective, hypers, config=None, cache=None, ctx=None, rng=None):
        super().__init__(objective, hypers, config, cache, ctx, rng)
        cache = cache or {}
        config = self.config
        self.model = cache.get("model") or getattr(Models, config["model"]["type"])(config["model"])
        self.score = cache.get("score") or getattr(Scores, config["score"]["type"])(config["score"])
        self.last = None

    def configure(self, config):
        """bots/bayesopt.lua:35-53."""
        config = super().configure(config)
        model = config.get("model") or {}
        for k, v in (("type", "gp_regressor"), ("kernel", "ardse"), ("nzModel", "GaussianNoise_iso"),
                     ("mean", "constant"), ("sampler", "slice")):
            model.setdefault(k, v)
        config["model"] = model
        return config

    def acquire(self, hyps=None, want_score=False):
        """Device body of bot:eval + bot:nominate: returns (score or None, idx_compacted, best, nan_count)."""
        X_obs, Y_obs = self.observed, self.responses[:, 0]
        kind, tradeoff, bound, sign = Scores.score_args(self.score)
        fmin = float(Y_obs.min())
        if self.model.class_() == "bot7.models.dngo":                # bots/bayesopt.lua:65-66
            raise NotImplementedError("dngo acquisition goes through models.dngo.predict + scores")
        if hyps is None:
            self.model.sample_hypers(X_obs, Y_obs)                   # :68 priming draw (result unused)
            nS = self.config["bot"]["nSamples"]
            hyps = np.stack([self.model.parse_hypers(self.model.sample_hypers(X_obs, Y_obs, None, None, True))[0]
                             for _ in range(nS)])                    # :73-75
        f = Models.GPFactors(X_obs, Y_obs, hyps, self.model.config["kernel"], self.model.config.get("noiseless", False),
                             L.FIT_PREDICT, self.ctx)
        M = self.grid.rows()
        score = np.empty(M) if want_score else None
        am, amo, best, nn = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
        try:
            L.check(L.lib().b7_acq_score(f.handle, self.grid.handle, kind, tradeoff, bound, sign, fmin,
                                         L.dptr(score) if want_score else None, C.byref(am), C.byref(amo), C.byref(best),
                                         C.byref(nn)), "b7_acq_score")
        finally:
            f.free()                                                 # multi-GB of factors: not left to the GC on an error
        self.last = {"argmax": am.value, "argmax_original": amo.value, "best": best.value, "nan_count": nn.value}
        if want_score:
            score = score[self._live_mask(M)]                        # compacted numbering, like the reference
        return score, am.value, best.value, nn.value

    def _live_mask(self, M):
        live = np.ones(M, dtype=bool)
        live[[r - 1 for r in self.removed_original]] = False
        return live

    def eval(self, candidates=None):
        """bots/bayesopt.lua:56-82 -> averaged score over the live candidates."""
        return self.acquire(want_score=True)[0]

    def nominate(self, candidates=None):
        """bots/bayesopt.lua:85-99."""
        if self.nTrials <= self.config["bot"]["nInitial"]:
            return int(self.rng.random() * self.grid.size()) + 1     # :90-91 floor(U*M)+1
        _, idx, _, nan_count = self.acquire()
        if idx == 0:
            raise RuntimeError("acquisition returned no finite score (%d NaN)" % nan_count)
        return idx
