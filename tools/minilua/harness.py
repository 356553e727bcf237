"""Builds an interpreter in which lua/bot7_b200/*.lua can run: torch + nn + ffi stand-ins, the `bot7` global with
stand-ins for the reference classes the glue subclasses or calls (the reference tree is Lua that cannot travel to the GPU
box and may not be copied: the fixtures below are minimal re-statements of the PROTOCOL -- constructor arguments and the
fields the glue reads -- not of the reference's code), and `require` resolving 'bot7_b200[.x]' to the glue directory."""
import ctypes as C
import io
import os

from . import ffi as ffi_mod
from . import torch7
from .interp import Interpreter, LuaError, LuaTable  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LUA_DIR = os.path.join(ROOT, "lua", "bot7_b200")
LIB_PATH = os.path.join(ROOT, "bot7_b200", "libbot7_b200.so")

# What `require 'bot7'` leaves behind as far as the glue is concerned (reference init.lua:28-40 creates the global table and
# its sub-tables; bots/abstract.lua:19-170 gives a bot its fields; samplers/slice.lua:24-28 is called as
# sampler(f, X0, opt, f_args) and returns an nSamples x dim tensor).  Stand-ins, written for the tests.
BOT7_FIXTURE = r"""
bot7 = {grids = {}, scores = {}, models = {}, bots = {}, samplers = {}, utils = {}}
do
  local g = torch.class('bot7.grids.abstract');   function g:__init() end
  function g:__call__(config) return self:generate(config or self.config) end     -- grids/abstract.lua:24-27
  local s = torch.class('bot7.scores.abstract');  function s:__init() end
  local m = torch.class('bot7.models.abstract');  function m:__init() end
  function m:cache() return {} end

  -- DNGO parent: owns the network; the glue reads self.network / self.basis / self.config and calls update_network
  local d, dp = torch.class('bot7.models.dngo', 'bot7.models.abstract')
  function d:__init(config, cache, X, Y)
    dp.__init(self)
    self.config  = config
    self.network = config.network
    self.basis   = config.basis
    self.updates = 0
  end
  function d:update_network(X0, Y0) self.updates = self.updates + 1 end

  -- bot parent: the fields bots/abstract.lua sets up and bots/bayesopt.lua reads
  local b = torch.class('bot7.bots.bayesopt')
  function b:__init(objective, hypers, config, cache)
    self.objective, self.hypers, self.config = objective, hypers, config
    self.candidates = config.candidates
    self.model, self.score = config.model, config.score
    self.observed, self.responses = config.observed, config.responses
    self.nTrials = config.nTrials or 0
  end

  -- slice sampler with the call protocol of samplers/slice.lua: univariate slices along random directions, step out, shrink
  local sl = torch.class('bot7.samplers.slice')
  function sl:__init() end
  function sl:__call__(f, X0, opt, f_args)
    opt = opt or {}
    local n, width, max_step = opt.nSamples or 1, opt.width or 1.0, opt.max_step or 50
    local dim = X0:size(2)
    local out = torch.Tensor(n, dim)
    local x   = X0[1]:clone()
    self.evals = self.evals or 0
    local function F(z) self.evals = self.evals + 1; return f(z:view(1, -1), f_args) end
    for i = 1, n do
      local dir = torch.randn(dim); dir:div(dir:norm())
      local f0  = F(x)
      assert(f0 > -math.huge, 'slice sampler: the chain starts at a point of zero density')
      local y   = f0 + math.log(1 - torch.rand(1)[1])
      local u   = torch.rand(1)[1]
      local lo, hi = -u * width, (1 - u) * width
      local steps = 0
      while steps < max_step and F(x + dir * lo) > y do lo = lo - width; steps = steps + 1 end
      steps = 0
      while steps < max_step and F(x + dir * hi) > y do hi = hi + width; steps = steps + 1 end
      local z
      for _ = 1, 200 do
        local t = lo + (hi - lo) * torch.rand(1)[1]
        z = x + dir * t
        if F(z) > y then break end
        if t < 0 then lo = t else hi = t end
        z = nil
      end
      x = z or x
      out[i]:copy(x)
    end
    return out
  end
end
"""


def default_lib_resolver(name):
    path = name if os.path.sep in name else LIB_PATH
    return C.CDLL(path)


class GlueRuntime:
    def __init__(self, seed=0, lib_resolver=default_lib_resolver, fixture=True, stdout=None):
        self.stdout = stdout if stdout is not None else io.StringIO()
        self.I = Interpreter(search_path=[("bot7_b200", LUA_DIR)], stdout=self.stdout)
        self.torch = torch7.install(self.I, seed)
        self.ffi = ffi_mod.Runtime(self.I, lib_resolver)
        self.ffi.install()
        if fixture:
            self.I.run(BOT7_FIXTURE, "=bot7 fixture")

    def run(self, text, name="=test"):
        return self.I.run(text, name)

    def require(self, name):
        return self.I.G.get("require")(name)

    def tensor(self, array, ttype="torch.DoubleTensor"):
        import numpy as np
        return torch7.Tensor(np.ascontiguousarray(np.array(array, dtype=torch7.TYPES[ttype])), ttype)

    def set_global(self, name, value):
        self.I.G.set(name, value)

    def close(self):
        # handles first (finalizers in reverse creation order), the context last: it was created first
        self.ffi.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
