-- bot7.scores.expected_improvement / confidence_bound: same constructors and call signature as
-- scores/expected_improvement.lua:27-41 and scores/confidence_bound.lua:29-42; compute() runs
-- the fused scoring kernel through b7_score_moments.
local B   = require('bot7_b200.ffi')
local ffi = require('ffi')

local function score_moments(kind, fval, fvar, fmin, tradeoff, bound, sign)
  local mean = fval:contiguous():double():view(-1)
  local var  = fvar:contiguous():double():view(-1)
  local M    = mean:nElement()
  local out  = torch.DoubleTensor(M)
  B.check(B.C.b7_score_moments(B.context(), kind, mean:data(), var:data(), 1, M, tradeoff, bound, sign,
                               fmin, out:data(), nil, nil, nil), 'b7_score_moments')
  return out
end

---------------- expected improvement
local EI, parent = torch.class('bot7_b200.scores.expected_improvement', 'bot7.scores.abstract')

function EI:__init(config)
  parent.__init(self)
  local config = config or {}
  config['tradeoff']   = config.tradeoff or 0.0
  config['nFantasies'] = config.nFantasies or 100
  self.config = config
end

function EI:__call__(model, hyp, X_obs, Y_obs, X_hid, X_pend, config)
  local hyp    = hyp or model.hyp
  local config = config or self.config
  return EI.eval(model, hyp, X_obs, Y_obs, X_hid, X_pend, config)
end

function EI.eval(model, hyp, X_obs, Y_obs, X_hid, X_pend, config)
  local pred  = model:predict(X_obs, Y_obs, X_hid, hyp, {mean=true, var=true})
  local fmins = Y_obs:min(1)
  return EI.compute(pred.mean, pred.var, fmins, config.tradeoff)
end

function EI.compute(fval, fvar, fmin, tradeoff)
  -- the reference reads an undefined global `config` here (scores/expected_improvement.lua:70)
  return score_moments(B.C.B7_SCORE_EI, fval, fvar, fmin:min(), tradeoff or 0.0, B.C.B7_BOUND_LOWER, -1.0)
end

---------------- confidence bound
local CB, parentCB = torch.class('bot7_b200.scores.confidence_bound', 'bot7.scores.abstract')

function CB:__init(config)
  local config = config or {}
  config['tradeoff']   = config.tradeoff or 1.0
  config['nFantasies'] = config.nFantasies or 100
  config['bound']      = config.bound or 'lower'
  config['sign']       = config.sign or -1.0
  self.config = config
end

function CB:__call__(model, hyp, X_obs, Y_obs, X_hid, X_pend, config)
  local hyp    = hyp or model.hyp
  local config = config or self.config
  local pred   = model:predict(X_obs, Y_obs, X_hid, hyp, {mean=true, var=true})
  return CB.compute(pred.mean, pred.var, config)
end

function CB.compute(fval, fvar, config)
  local bound = (config.bound:lower() == 'upper') and B.C.B7_BOUND_UPPER or B.C.B7_BOUND_LOWER
  return score_moments(B.C.B7_SCORE_CB, fval, fvar, 0.0, config.tradeoff or 1.0, bound, config.sign or -1.0)
end

return { expected_improvement = EI, confidence_bound = CB }
