-- Speculative slice sampler: the chain of bot7.samplers.slice (samplers/slice.lua:51-168), fewer sequential density evaluations.
--
-- The reference sampler calls f strictly one point at a time, and for GP hyper-parameters every call is a device fit whose cost at
-- small and medium N is latency, not arithmetic (one resident factor: 0.24 ms at N = 512, 2.0 ms at N = 4096; eight factors in one
-- call: 0.28 ms / 6.4 ms).  Here f_batch(points [k x h], f_args) -> k values evaluates several points in one device call
-- (model:log_density_batch = one b7_gp_refit of k factors), and the control flow asks for points BEFORE it knows it needs them:
--   * the current point, the first right and left bracket ends (their positions depend on the RNG only) and -- betting that no
--     stepping out is needed -- the first shrink proposals go out as one batch;
--   * stepping out evaluates the next `width` ends of the side being extended at once (stepping out draws no random numbers);
--   * stepping in pre-computes the next `width` proposals under the assumption that each one is rejected (the bracket update after
--     a rejection depends only on the sign of the proposal).
-- Speculative results the sequential algorithm would not have requested are discarded and the generator is rewound
-- (torch.getRNGState / setRNGState) to exactly the state the sequential algorithm would have left it in, so for a given generator
-- state the samples are those of samplers/slice.lua, bit for bit.  Same algorithm as bot7_b200/samplers.py:slice_speculative;
-- tests/test_lua_exec.py runs it next to the EXECUTED reference sampler (CPU) and next to the Python twin on the GPU.
require('bot7_b200.ffi')          -- creates the package table bot7_b200.samplers that torch.class stores the class in
local sampler = torch.class('bot7_b200.samplers.slice_speculative')

function sampler:__init()
  self.calls, self.evals = 0, 0
end

-- samplers/slice.lua:32-48
local function configure(opt)
  local opt = opt or {}
  opt['max_step'] = opt.max_step or 1e3
  opt['nSamples'] = opt.nSamples or 1
  if opt.step_out ~= false then opt['step_out'] = true end
  if opt.logspace ~= false then opt['logspace'] = true end
  return opt
end

-- a rejected proposal p replaces the bracket end on its side, per coordinate (:153-161)
local function shrink(left, right, p)
  local pos, neg = p:gt(0), p:lt(0)
  if pos:any() then right = right:clone():maskedCopy(pos, p:maskedSelect(pos)) end
  if neg:any() then left  = left:clone():maskedCopy(neg, p:maskedSelect(neg)) end
  return left, right
end

-- `width` shrink proposals under the assumption that each earlier one is rejected; the RNG state after every draw is kept so
-- that the caller can rewind to where the sequential algorithm would be
local function propose(left, right, width)
  local l, r, props, states = left, right, {}, {}
  for _ = 1, width do
    local u = torch.rand(1)[1]
    states[#states + 1] = torch.getRNGState()
    local p = l + (r - l) * u                                   -- :136
    props[#props + 1] = {p = p, l = l, r = r}
    if p:eq(0.0):any() then break end
    l, r = shrink(l, r, p)
  end
  return props, states
end

function sampler:evaluate(f_batch, f_args, x0, dir, dxs)
  local k = #dxs
  local pts = torch.Tensor(k, x0:size(2))
  for i = 1, k do pts[i]:copy(x0 + dir:clone():cmul(dxs[i])) end     -- the point f_dx(dx) evaluates (:100-103)
  self.calls, self.evals = self.calls + 1, self.evals + k
  local v, out = f_batch(pts, f_args), {}
  for i = 1, k do out[i] = v[i] end
  return out
end

function sampler:directed_slice(opt, f_batch, f_args, dir, x0, width)
  local xDim     = x0:size(2)
  local stepsize = opt.widths or torch.Tensor(1, xDim):fill(opt.width or 1.0)
  local zero     = torch.zeros(1, xDim)
  -- RNG order of the reference: the slice level (:108), then the bracket (:114)
  local u
  if opt.logspace then u = torch.log(torch.rand(1))[1] else u = torch.rand(1)[1] end
  local right = torch.rand(1, xDim):cmul(stepsize)
  local left  = right - stepsize
  -- first device call: the point itself, both bracket ends and the first proposals for this bracket
  local state0 = torch.getRNGState()
  local props, states = propose(left, right, math.max(width - 3, 1))
  local dxs = {zero, right, left}
  for i = 1, #props do dxs[#dxs + 1] = props[i].p end
  local vals = self:evaluate(f_batch, f_args, x0, dir, dxs)
  local f0, fr, fl, ys = vals[1], vals[2], vals[3], {}
  for i = 4, #vals do ys[#ys + 1] = vals[i] end
  local Y
  if opt.logspace then Y = f0 + u else Y = f0 * u end
  local pending = {props = props, states = states, ys = ys}
  if opt.step_out and (fr > Y or fl > Y) then
    pending = nil                                               -- bet lost: discard the proposals, rewind the generator
    torch.setRNGState(state0)
    for _, side in ipairs{1, -1} do
      local stop, fe, itr, cache = right, fr, 0, {}
      if side < 0 then stop, fe = left, fl end
      while fe > Y and itr < opt.max_step do                    -- :118-130
        itr = itr + 1
        if side > 0 then stop = stop + stepsize else stop = stop - stepsize end
        if #cache == 0 then
          local ends, e = {}, stop
          for _ = 1, width do                                   -- repeated addition, exactly like the sequential loop
            ends[#ends + 1] = e
            if side > 0 then e = e + stepsize else e = e - stepsize end
          end
          local fv = self:evaluate(f_batch, f_args, x0, dir, ends)
          for i = 1, #ends do cache[i] = {ends[i], fv[i]} end
        end
        local c = table.remove(cache, 1)
        stop, fe = c[1], c[2]
      end
      if side > 0 then right = stop else left = stop end
    end
  end
  -- stepping in (:134-164)
  local dx = zero
  while true do
    local ps, sts, fs
    if pending then
      ps, sts, fs = pending.props, pending.states, pending.ys
      pending = nil
    else
      ps, sts = propose(left, right, width)
      local pts = {}
      for i = 1, #ps do pts[i] = ps[i].p end
      fs = self:evaluate(f_batch, f_args, x0, dir, pts)
    end
    local done = false
    for i = 1, #ps do
      local y = fs[i]
      dx, left, right = ps[i].p, ps[i].l, ps[i].r
      torch.setRNGState(sts[i])                                 -- the generator exactly as after this proposal's draw
      if y ~= y then
        print('Error: samplers.slice encountered a NaN')
        done = true
        break
      end
      if y > Y then done = true; break end
      if dx:eq(0.0):any() then
        print('Error: samplers.slice shrank to zero')
        done = true
        break
      end
      left, right = shrink(left, right, dx)
    end
    if done then break end
  end
  return x0 + dir:clone():cmul(dx)
end

-- sampler(f_batch, X0, opt, f_args, width) -> nSamples x dim tensor (samplers/slice.lua:24-28,51-89 without the Gibbs sweep)
function sampler:__call__(f_batch, X0, opt, f_args, width)
  local opt   = configure(opt)
  local width = math.max(width or 4, 1)
  assert(not opt.gibbs, 'bot7_b200.samplers.slice_speculative: Gibbs sweeps use bot7.samplers.slice')
  local X0 = X0:clone():repeatTensor(opt.nSamples, 1)
  local N, xDim = X0:size(1), X0:size(2)
  local samples = torch.Tensor(N, xDim)
  for n = 1, N do
    local x0  = X0[{{n}, {}}]
    local dir = torch.randn(1, xDim)
    dir = dir:div(dir:norm())
    samples[n] = self:directed_slice(opt, f_batch, f_args, dir, x0, width)
  end
  return samples
end

return sampler
