-- bot7.models.dngo with the hand-off of models/dngo.lua:155-174 routed to the GPU: the network update (trainer,
-- :126-153) and every other method stay the parent's; what changes is how Z1 = basis(X_hid) and
-- predictor:predict(Z0, Y0, Z1, nil, hyp, req) are evaluated for the candidate grid:
--   * the Linear / ReLU stack below the basis layer is read out of self.network and applied to the device grid by
--     b7_mlp_features (no 32-row minibatch loop, no host copy of Z1), the head is b7_blr_fit + b7_blr_predict;
--   * bots.bayesopt (bayesopt.lua) calls model:acquire(...), which fuses basis + head + score + argmax in one pass
--     over the grid (b7_dngo_score) and never stores Z1.
-- BLR hyper rows [log alpha_p, log beta, m] follow oracle/SPEC.md (gpTorch7's bayes_linear is not available).
local B   = require('bot7_b200.ffi')
local ffi = require('ffi')

local dngo, parent = torch.class('bot7_b200.models.dngo', 'bot7.models.dngo')

function dngo:__init(config, cache, X, Y)
  parent.__init(self, config, cache, X, Y)
end

function dngo:class() return 'bot7.models.dngo' end

-- weights of the nn.Linear modules up to (and including) the basis layer: arrays for b7_mlp_features / b7_dngo_score
function dngo:basis_stack()
  local Ws, bs, dims, relu_last = {}, {}, {}, false
  for idx = 1, self.network:size() do
    local mod = self.network:get(idx)
    if mod.weight then
      assert(torch.type(mod) == 'nn.Linear', 'bot7_b200.models.dngo: the basis must be a Linear/ReLU stack (nnTools/builder.lua:133-159)')
      Ws[#Ws + 1] = mod.weight:contiguous():double()         -- h_out x h_in, like torch nn.Linear.weight
      bs[#bs + 1] = mod.bias:contiguous():double()
      if #dims == 0 then dims[1] = mod.weight:size(2) end
      dims[#dims + 1] = mod.weight:size(1)
      relu_last = false
    elseif torch.type(mod) == 'nn.ReLU' then
      relu_last = true
    end
    if mod == self.basis then break end
  end
  local n   = #Ws
  local cd  = ffi.new('int[?]', n + 1, dims)
  local cW  = ffi.new('const double*[?]', n)
  local cb  = ffi.new('const double*[?]', n)
  for l = 1, n do cW[l - 1] = Ws[l]:data(); cb[l - 1] = bs[l]:data() end
  return {n = n, dims = cd, W = cW, b = cb, relu_last = relu_last and 1 or 0, keep = {Ws, bs}, zDim = dims[#dims]}
end

-- Z = basis(X) on the host for the (few) observations, through the network itself (models/dngo.lua:155-162)
function dngo:host_basis(X)
  local bsz = self.config.update.schedule.batchsize
  local N   = X:size(1)
  local Z   = torch.DoubleTensor(N, self.config.zDim)
  self.network:evaluate()
  for tail = 1, N, bsz do
    local head = math.min(tail + bsz - 1, N)
    self.network:forward(X:sub(tail, head))
    Z:sub(tail, head):copy(self.basis.output)
  end
  return Z
end

function dngo:blr_hyp(Y0, hyp)
  if torch.isTensor(hyp) then return hyp:contiguous():double() end
  local p = self.config.predictor or {}
  return torch.DoubleTensor{{math.log(p.alpha_p or 1.0), math.log(p.beta or 1e2), Y0:mean()}}
end

function dngo:blr_fit(Z0, Y0, hyp)
  local Z0, Y0, hyp = Z0:contiguous():double(), Y0:contiguous():double(), self:blr_hyp(Y0, hyp)
  local box, info = ffi.new('b7_blr*[1]'), ffi.new('int[?]', hyp:size(1))
  B.check(B.C.b7_blr_fit(B.context(), Z0:data(), Y0:data(), Z0:size(1), Z0:size(2), hyp:data(), hyp:size(1), box, info), 'b7_blr_fit')
  return ffi.gc(box[0], B.C.b7_blr_free), hyp:size(1)
end

---------------- models/dngo.lua:108-175: same signature and return value
function dngo:predict(X0, Y0, X1, hyp, req, skip)
  if not skip then parent.update_network(self, X0, Y0) end       -- the network update of :126-153 (see install())
  local Z0      = self:host_basis(X0)
  local blr, S  = self:blr_fit(Z0, Y0, hyp)
  local st      = self:basis_stack()
  local X1      = X1:contiguous():double()
  local M       = X1:size(1)
  local gbox, fbox = ffi.new('b7_grid*[1]'), ffi.new('b7_grid*[1]')
  B.check(B.C.b7_grid_from_host(B.context(), X1:data(), M, X1:size(2), gbox), 'b7_grid_from_host')
  local grid = ffi.gc(gbox[0], B.C.b7_grid_free)
  B.check(B.C.b7_mlp_features(B.context(), grid, st.n, st.dims, st.W, st.b, st.relu_last, fbox), 'b7_mlp_features')
  local feats = ffi.gc(fbox[0], B.C.b7_grid_free)
  local Z1 = torch.DoubleTensor(M, st.zDim)
  B.check(B.C.b7_grid_read(feats, 0, M, Z1:data()), 'b7_grid_read')
  local mean, var = torch.zeros(M, 1), torch.zeros(M, 1)
  local m_s, v_s  = torch.DoubleTensor(M, 1), torch.DoubleTensor(M, 1)
  for s = 0, S - 1 do                                            -- 'marginalize': average over the (alpha_p, beta) rows
    B.check(B.C.b7_blr_predict(blr, s, Z1:data(), M, m_s:data(), v_s:data()), 'b7_blr_predict')
    mean:add(m_s); var:add(v_s)
  end
  collectgarbage()
  return {mean = mean:div(S), var = var:div(S)}
end

---------------- device body of bot:eval + bot:nominate for DNGO (bots/bayesopt.lua:65-66,96): one pass over the grid
function dngo:acquire(X0, Y0, grid_dev, kind, tradeoff, bound, sign, skip)
  if not skip then parent.update_network(self, X0, Y0) end
  local blr = self:blr_fit(self:host_basis(X0), Y0, nil)
  local st  = self:basis_stack()
  local argmax, orig = ffi.new('int64_t[1]'), ffi.new('int64_t[1]')
  local best, nans   = ffi.new('double[1]'), ffi.new('int64_t[1]')
  B.check(B.C.b7_dngo_score(blr, grid_dev, st.n, st.dims, st.W, st.b, st.relu_last, kind, tradeoff, bound, sign, Y0:min(),
                            nil, argmax, orig, best, nans), 'b7_dngo_score')
  return tonumber(argmax[0]), best[0], tonumber(nans[0])
end

return dngo
