--[[
bot7_b200: drop-in overrides for the surrogate-fit-and-acquisition path of bot7.

Usage (after `require 'bot7'`):
    require('bot7_b200').install()
which replaces, in the global `bot7` table (reference init.lua:28-40),
    bot7.grids.sobol          -> grids_sobol.lua   (grid:generate on the GPU, bit-exact)
    bot7.scores.*             -> scores.lua        (EI.compute / conf_bound.compute fused on the GPU)
    bot7.models.gp_regressor  -> models_gp.lua     (predict / log-density through b7_gp_fit/predict)
    bot7.bots.bayesopt        -> bayesopt.lua      (eval+nominate = one batched device call)
Everything else in bot7 (config tables, trial loop, objectives, nnTools) is untouched.
No Lua runtime exists in the build image; this glue is exercised through its Python twin
(bot7_b200/*.py), which calls the same C symbols with the same arguments.
--]]
local M = {}
M.ffi      = require('bot7_b200.ffi')
M.grids    = { sobol = require('bot7_b200.grids_sobol') }
M.scores   = require('bot7_b200.scores')
M.models   = { gp_regressor = require('bot7_b200.models_gp') }
M.bots     = { bayesopt = require('bot7_b200.bayesopt') }

function M.install()
  assert(bot7, "require 'bot7' first")
  bot7.grids.sobol                  = M.grids.sobol
  bot7.scores.expected_improvement  = M.scores.expected_improvement
  bot7.scores.confidence_bound      = M.scores.confidence_bound
  bot7.models.gp_regressor          = M.models.gp_regressor
  bot7.bots.bayesopt                = M.bots.bayesopt
  return bot7
end

return M
