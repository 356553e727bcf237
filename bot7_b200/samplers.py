"""bot7.samplers.slice -- host control flow of reference samplers/slice.lua (SURVEY a-15).

Strictly sequential; every call of `f` is one device evaluation (for GP hyper-parameters: one
b7_gp_fit with B7_FIT_LOGML_ONLY).  The RNG is host state (torch MT19937 in the reference, numpy
here) and is not reproduced; the control flow is.
"""
from __future__ import annotations

import numpy as np


class slice:  # noqa: A001  (name kept from the reference)
    def __call__(self, f, X0, opt=None, f_args=None, rng=None):
        opt = self.configure(opt)                                   # samplers/slice.lua:26
        return self.sample(f, X0, opt, f_args, rng or np.random.default_rng())

    @staticmethod
    def configure(opt=None):
        """samplers/slice.lua:32-48."""
        opt = dict(opt or {})
        opt["max_step"] = opt.get("max_step") or 1e3
        opt["nSamples"] = opt.get("nSamples") or 1
        if opt.get("step_out") is not False:
            opt["step_out"] = True
        if opt.get("logspace") is not False:
            opt["logspace"] = True
        return opt

    @staticmethod
    def sample(f, X0, opt, f_args, rng):
        """samplers/slice.lua:51-89."""
        X0 = np.atleast_2d(np.asarray(X0, dtype=np.float64))
        X0 = np.tile(X0.copy(), (int(opt["nSamples"]), 1))
        N, xDim = X0.shape
        samples = np.empty((N, xDim))
        if opt.get("gibbs"):
            for n in range(N):
                x0 = X0[n:n + 1]
                x1 = np.zeros((1, xDim))
                d_vec = np.zeros((1, xDim))
                order = rng.permutation(xDim)
                for itr in range(xDim):
                    d = order[itr]
                    d_vec[0, d] = 1.0
                    x1[0, d] = slice.directed_slice(opt, f, f_args, d_vec, x0, rng)[0, d]
                    d_vec[0, d] = 0.0
                samples[n] = x1
        else:
            for n in range(N):
                x0 = X0[n:n + 1]
                d_vec = rng.standard_normal((1, xDim))
                d_vec = d_vec / np.linalg.norm(d_vec)
                samples[n] = slice.directed_slice(opt, f, f_args, d_vec, x0, rng)
        return samples

    @staticmethod
    def directed_slice(opt, f, f_args, d_vec, x0, rng):
        """samplers/slice.lua:92-168."""
        xDim = x0.shape[1]
        stepsize = opt.get("widths")
        if stepsize is None:
            stepsize = np.full((1, xDim), opt.get("width") or 1.0)   # :95

        def f_dx(dx=None):                                            # :100-103
            dx = np.zeros((1, xDim)) if dx is None else dx
            return f(x0 + d_vec * dx, f_args)

        Y = f_dx()                                                    # :106
        if opt["logspace"]:
            Y = Y + np.log(rng.random())                              # :108
        else:
            Y = Y * rng.random()
        right = rng.random((1, xDim)) * stepsize                      # :114
        left = right - stepsize                                       # :115
        if opt["step_out"]:                                           # :118-130
            itr = 0
            while f_dx(right) > Y and itr < opt["max_step"]:
                itr += 1
                right = right + stepsize
            itr = 0
            while f_dx(left) > Y and itr < opt["max_step"]:
                itr += 1
                left = left - stepsize
        dx = np.zeros((1, xDim))
        while True:                                                   # :134-164
            dx = left + (right - left) * rng.random()
            y = f_dx(dx)
            if y != y:                                                # :139-142
                print("Error: samplers.slice encountered a NaN")
                break
            if y > Y:                                                 # :144
                break
            if (dx == 0.0).any():                                     # :148-151
                print("Error: samplers.slice shrank to zero")
                break
            pos = dx > 0                                              # :153-161
            right = np.where(pos, dx, right)
            neg = dx < 0
            left = np.where(neg, dx, left)
        return x0 + d_vec * dx
