"""bot7.scores -- acquisition functions (host-side mirror of reference scores/*.lua).

`score(model, hyp, X_obs, Y_obs, X_hid, X_pend, config)` keeps the reference signature
(scores/expected_improvement.lua:35, scores/confidence_bound.lua:38) and returns an M-vector.
The static `compute` is EI.compute / conf_bound.compute on given moments.  All arithmetic runs in
the fused scoring kernel (csrc/score.cu) through b7_score_moments / b7_acq_score.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _dflt(c, key, value):
    # Lua's `x or default`: only nil/false fall through (0 is truthy in Lua)
    if c.get(key) is None or c.get(key) is False:
        c[key] = value


def _score_moments(kind, mean, var, fmin, tradeoff, bound, sign, ctx=None):
    ctx = ctx or L.Context.default()
    mean = np.atleast_2d(L.as_f64(mean))
    var = np.atleast_2d(L.as_f64(var))
    S, M = mean.shape
    score = np.empty(M)
    am, best, nn = C.c_int64(), C.c_double(), C.c_int64()
    L.check(L.lib().b7_score_moments(ctx.handle, kind, L.dptr(mean), L.dptr(var), S, M, float(tradeoff), int(bound),
                                     float(sign), float(fmin), L.dptr(score), C.byref(am), C.byref(best), C.byref(nn)),
            "b7_score_moments")
    return score, am.value, best.value, nn.value


class _abstract:
    kind = None

    def __call__(self, model, hyp, X_obs, Y_obs, X_hid, X_pend=None, config=None):
        hyp = model.hyp if hyp is None else hyp
        config = self.config if config is None else config
        return type(self).eval(model, hyp, X_obs, Y_obs, X_hid, X_pend, config)

    @classmethod
    def eval(cls, model, hyp, X_obs, Y_obs, X_hid, X_pend, config):
        X_obs = np.atleast_2d(L.as_f64(X_obs))
        X_hid = np.atleast_2d(L.as_f64(X_hid))
        Y_obs = L.as_f64(Y_obs).reshape(X_obs.shape[0], -1)
        if X_pend is not None and np.size(X_pend) > 0:
            # fantasy branch (scores/expected_improvement.lua:51-60): dead in the driven path
            # (bots/bayesopt.lua:66,76 never pass X_pend); kept for API completeness.
            X_pend = np.atleast_2d(L.as_f64(X_pend))
            nF = int(config.get("nFantasies", 100))
            Y_pend = model.fantasize(nF, X_obs, Y_obs[:, 0], X_pend, hyp)
            X_all = np.concatenate([X_obs, X_pend], 0)
            cols = []
            for fcol in range(nF):
                Y_all = np.concatenate([Y_obs[:, 0], Y_pend[:, fcol]])
                pred = model.predict(X_all, Y_all, X_hid, hyp, {"mean": True, "var": True})
                cols.append(cls._compute_vec(pred["mean"][:, 0], pred["var"][:, 0], float(Y_all.min()), config))
            return np.mean(np.stack(cols, 1), axis=1)               # :83-85 ei:mean(2)
        pred = model.predict(X_obs, Y_obs[:, 0], X_hid, hyp, {"mean": True, "var": True})   # :63
        fmin = float(Y_obs.min())                                                           # :64
        return cls._compute_vec(pred["mean"][:, 0], pred["var"][:, 0], fmin, config)


class expected_improvement(_abstract):
    kind = L.SCORE_EI

    def __init__(self, config=None):
        c = dict(config or {})
        _dflt(c, "tradeoff", 0.0)        # scores/expected_improvement.lua:30
        _dflt(c, "nFantasies", 100)      # :31
        self.config = c

    @staticmethod
    def compute(fval, fvar, fmin, tradeoff=0.0):
        """EI.compute (scores/expected_improvement.lua:69-88); `tradeoff` is explicit (the reference reads
        an undefined global `config`, SURVEY a-6)."""
        return _score_moments(L.SCORE_EI, np.reshape(fval, (1, -1)), np.reshape(fvar, (1, -1)), float(np.min(fmin)),
                              tradeoff, L.BOUND_LOWER, -1.0)[0]

    @classmethod
    def _compute_vec(cls, mean, var, fmin, config):
        return cls.compute(mean, var, fmin, config.get("tradeoff", 0.0))


class confidence_bound(_abstract):
    kind = L.SCORE_CB

    def __init__(self, config=None):
        c = dict(config or {})
        _dflt(c, "tradeoff", 1.0)        # scores/confidence_bound.lua:31
        _dflt(c, "nFantasies", 100)
        _dflt(c, "bound", "lower")       # :33
        _dflt(c, "sign", -1.0)           # :34
        self.config = c

    @staticmethod
    def compute(fval, fvar, config):
        """conf_bound.compute (scores/confidence_bound.lua:70-94)."""
        bound = (config.get("bound") or "lower").lower()
        if bound not in ("lower", "upper"):
            raise ValueError("bound must be 'lower' or 'upper'")
        tradeoff = 1.0 if config.get("tradeoff") is None else config["tradeoff"]
        sign = -1.0 if config.get("sign") is None else config["sign"]
        return _score_moments(L.SCORE_CB, np.reshape(fval, (1, -1)), np.reshape(fvar, (1, -1)), 0.0, tradeoff,
                              L.BOUND_LOWER if bound == "lower" else L.BOUND_UPPER, sign)[0]

    @classmethod
    def _compute_vec(cls, mean, var, fmin, config):
        return cls.compute(mean, var, config)

    @staticmethod
    def UCB(fval, fvar, tradeoff=1.0):
        return confidence_bound.compute(fval, fvar, {"tradeoff": tradeoff, "bound": "upper", "sign": 1.0})

    @staticmethod
    def LCB(fval, fvar, tradeoff=1.0):
        return confidence_bound.compute(fval, fvar, {"tradeoff": tradeoff, "bound": "lower", "sign": 1.0})


def score_args(score):
    """(kind, tradeoff, bound, sign) of a score object for the fused device path."""
    c = score.config
    if isinstance(score, expected_improvement):
        return L.SCORE_EI, float(c["tradeoff"]), L.BOUND_LOWER, -1.0
    bound = L.BOUND_LOWER if str(c["bound"]).lower() == "lower" else L.BOUND_UPPER
    return L.SCORE_CB, float(c["tradeoff"]), bound, float(c["sign"])
