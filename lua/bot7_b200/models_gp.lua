-- gp_regressor protocol (predict / log-density for sample_hypers) over b7_gp_fit / b7_gp_predict.
-- hyp layout: [log l_1..log l_d, log sigma_f, log sigma_n, m]  (oracle/SPEC.md).
local B   = require('bot7_b200.ffi')
local ffi = require('ffi')

local kernels = { ardse = B.C.B7_KERNEL_ARDSE, matern52 = B.C.B7_KERNEL_MATERN52 }
local model = torch.class('bot7_b200.models.gp_regressor', 'bot7.models.abstract')

function model:__init(config)
  self.config = config or {}
  self.config.kernel = self.config.kernel or 'ardse'
  self.hyp = nil
end

-- S x H hyper draws -> device factors (handle freed by the GC)
function model:fit(X, Y, hyp, flags)
  local X, Y, hyp = X:contiguous():double(), Y:contiguous():double(), hyp:contiguous():double()
  if hyp:dim() == 1 then hyp = hyp:view(1, -1) end
  local S, H = hyp:size(1), hyp:size(2)
  local box   = ffi.new('b7_gp*[1]')
  local info  = ffi.new('int[?]', S)
  local logml = torch.DoubleTensor(S)
  local jit   = torch.DoubleTensor(S)
  B.check(B.C.b7_gp_fit(B.context(), kernels[self.config.kernel], X:data(), Y:data(), X:size(1), X:size(2),
                        hyp:data(), S, H, self.config.noiseless and 1 or 0, flags or B.C.B7_FIT_PREDICT,
                        box, info, logml:data(), jit:data()), 'b7_gp_fit')
  for s = 1, S do
    if jit[s] > 0 then -- utils/math.lua:204-215
      print(string.format('Warning: utils.math.chol succeeded in factorizing the\ninput matrix after applying a jitter of %.2e', jit[s]))
    end
  end
  return ffi.gc(box[0], B.C.b7_gp_free), logml
end

function model:predict(X0, Y0, X1, hyp, req)
  local gp   = self:fit(X0, Y0, hyp or self.hyp)
  local X1   = X1:contiguous():double()
  local M    = X1:size(1)
  local mean = torch.DoubleTensor(M, 1)
  local var  = torch.DoubleTensor(M, 1)
  B.check(B.C.b7_gp_predict(gp, 0, X1:data(), M, mean:data(), var:data()), 'b7_gp_predict')
  return {mean = mean, var = var}
end

-- log p(y | hyp): the density the slice sampler evaluates (samplers/slice.lua:100-103).
-- X and y stay resident between evaluations: the first call fits, later calls b7_gp_refit.
function model:log_density(hyp, X, Y)
  local hyp = hyp:contiguous():double():view(1, -1)
  if self._density == nil or self._density_X ~= X or self._density_Y ~= Y then
    local gp, logml = self:fit(X, Y, hyp, B.C.B7_FIT_LOGML_ONLY)
    self._density, self._density_X, self._density_Y = gp, X, Y
    return logml[1]
  end
  local info, logml, jit = ffi.new('int[1]'), torch.DoubleTensor(1), torch.DoubleTensor(1)
  B.check(B.C.b7_gp_refit(self._density, hyp:data(), B.C.B7_FIT_LOGML_ONLY, info, logml:data(), jit:data()), 'b7_gp_refit')
  return logml[1]
end

function model:parse_hypers(h) return h end

return model
