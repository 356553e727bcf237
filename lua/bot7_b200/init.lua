--[[
bot7_b200: drop-in overrides for the surrogate-fit-and-acquisition path of bot7.

Usage (after `require 'bot7'`):
    require('bot7_b200').install()
which replaces, in the global `bot7` table (reference init.lua:28-40),
    bot7.grids.sobol          -> grids_sobol.lua   (grid:generate on the GPU, bit-exact)
    bot7.scores.*             -> scores.lua        (EI.compute / conf_bound.compute fused on the GPU)
    bot7.models.gp_regressor  -> models_gp.lua     (init / predict / sample_hypers / fantasize through b7_gp_fit, b7_gp_refit, b7_gp_predict)
    bot7.models.dngo          -> models_dngo.lua   (basis + BLR head on the GPU: b7_mlp_features, b7_blr_*, b7_dngo_score)
    bot7.bots.bayesopt        -> bayesopt.lua      (eval+nominate = one batched device call; config.bot.nGPU > 1 -> b7_comm_*)
Everything else in bot7 (config tables, trial loop, objectives, nnTools) is untouched.
No Lua runtime exists in the build image: this glue is executed by tests/test_lua_exec.py under the interpreter of
tools/minilua (Torch7 / LuaJIT-FFI stand-ins over the real library) and mirrored by its Python twin (bot7_b200/*.py).
--]]
local M = {}
M.ffi      = require('bot7_b200.ffi')
M.grids    = { sobol = require('bot7_b200.grids_sobol') }
M.scores   = require('bot7_b200.scores')
M.models   = { gp_regressor = require('bot7_b200.models_gp'), dngo = require('bot7_b200.models_dngo') }
M.bots     = { bayesopt = require('bot7_b200.bayesopt') }

function M.install()
  assert(bot7, "require 'bot7' first")
  bot7.grids.sobol                  = M.grids.sobol
  bot7.scores.expected_improvement  = M.scores.expected_improvement
  bot7.scores.confidence_bound      = M.scores.confidence_bound
  bot7.models.gp_regressor          = M.models.gp_regressor
  -- models/dngo.lua:126-153 (the network update inside predict) becomes a method the override can call on its own
  local ref_dngo = bot7.models.dngo
  if ref_dngo.update_network == nil then
    function ref_dngo:update_network(X0, Y0)
      self.state.dfdx:fill(0.0)
      local cache = {optimizer = self.optimizer, buffers = self.buffers, criterion = self.criterion, state = self.state}
      require('bot7.nnTools.trainer')(self.network, {xr = X0, yr = Y0}, self.config.update, cache)
    end
  end
  bot7.models.dngo                  = M.models.dngo
  bot7.bots.bayesopt                = M.bots.bayesopt
  return bot7
end

return M
