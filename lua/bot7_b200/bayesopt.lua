-- bots.bayesopt with the body of eval/nominate (bots/bayesopt.lua:56-99) routed to one batched
-- device call: S draws -> S factors (b7_gp_fit) -> b7_acq_score over the device-resident grid.
local B   = require('bot7_b200.ffi')
local ffi = require('ffi')

local title  = 'bot7_b200.bots.bayesopt'
local parent = 'bot7.bots.bayesopt'
local bot, parent = torch.class(title, parent)

function bot:__init(objective, hypers, config, cache)
  parent.__init(self, objective, hypers, config, cache)
  -- keep the grid on the device as well; self.candidates stays for callers that read it
  local box = ffi.new('b7_grid*[1]')
  local X = self.candidates:contiguous():double()
  B.check(B.C.b7_grid_from_host(B.context(), X:data(), X:size(1), X:size(2), box), 'b7_grid_from_host')
  self.grid_dev = ffi.gc(box[0], B.C.b7_grid_free)
end

function bot:nominate(candidates)
  if self.nTrials <= self.config.bot.nInitial then
    local idx = torch.rand(1):mul(self.candidates:size(1)):long():add(1)
    B.check(B.C.b7_grid_remove(self.grid_dev, idx[1], nil), 'b7_grid_remove')
    return idx
  end
  local X_obs, Y_obs = self.observed, self.responses
  local nSamples = self.config.bot.nSamples
  self.model:sample_hypers(X_obs, Y_obs)                        -- bots/bayesopt.lua:68
  local hyps = {}
  for s = 1, nSamples do                                        -- :73-75
    hyps[s] = self.model:parse_hypers(self.model:sample_hypers(X_obs, Y_obs, nil, nil, true)):view(1, -1)
  end
  local gp = self.model:fit(X_obs, Y_obs, torch.cat(hyps, 1))
  local sc = self.score.config
  local kind = (torch.type(self.score):find('confidence_bound')) and B.C.B7_SCORE_CB or B.C.B7_SCORE_EI
  local bound = (sc.bound and sc.bound:lower() == 'upper') and B.C.B7_BOUND_UPPER or B.C.B7_BOUND_LOWER
  local argmax, orig = ffi.new('int64_t[1]'), ffi.new('int64_t[1]')
  local best, nans   = ffi.new('double[1]'), ffi.new('int64_t[1]')
  B.check(B.C.b7_acq_score(gp, self.grid_dev, kind, sc.tradeoff, bound, sc.sign or -1.0, Y_obs:min(),
                           nil, argmax, orig, best, nans), 'b7_acq_score')
  local idx = torch.LongTensor{tonumber(argmax[0])}             -- compacted numbering, as score:max(1)
  B.check(B.C.b7_grid_remove(self.grid_dev, idx[1], nil), 'b7_grid_remove')
  return idx
end

return bot
