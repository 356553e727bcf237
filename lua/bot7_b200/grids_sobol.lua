-- bot7.grids.sobol over b7_sobol_generate (replaces grids/sobol.lua:58-90,216-335).
local B   = require('bot7_b200.ffi')
local ffi = require('ffi')

local title  = 'bot7_b200.grids.sobol'
local parent = 'bot7.grids.abstract'
local grid, parent = torch.class(title, parent)

function grid:__init(config)
  parent.__init(self)
  local C = config or {}
  C.max_dims   = C.max_dims or 40
  C.log_max    = C.log_max or 30
  C.bit_precis = C.bit_precis or 32
  assert(C.size)
  assert(C.dims and C.dims < C.max_dims)
  self.config = C
end

-- host tensor, exactly like the reference (size x dims DoubleTensor)
function grid:generate(config)
  local config = config or self.config
  local skip   = config.skip or 1
  local out    = torch.DoubleTensor(config.size, config.dims)
  local both   = config.mins and config.maxes
  local mins   = both and config.mins:contiguous():double() or nil
  local maxes  = both and config.maxes:contiguous():double() or nil
  B.check(B.C.b7_sobol_generate(B.context(), config.dims, skip, config.size,
                                B.ptr(mins), B.ptr(maxes), out:data(), nil), 'b7_sobol_generate')
  if not both then -- one-sided variants, grids/sobol.lua:82-86
    if config.mins then
      out:add(torch.add(config.mins, out:min(1)[1]):expandAs(out))
    elseif config.maxes then
      out:cmul(torch.cdiv(config.maxes, out:max(1)[1]):expandAs(out))
    end
  end
  return out
end

-- device-resident grid handle for bots.bayesopt (first/count select this rank's shard)
function grid:generate_device(config, first, count)
  local config = config or self.config
  local skip   = config.skip or 1
  first = first or 0
  count = count or (config.size - first)
  local box = ffi.new('b7_grid*[1]')
  B.check(B.C.b7_sobol_generate(B.context(), config.dims, skip + first, count,
                                B.ptr(config.mins and config.mins:contiguous()),
                                B.ptr(config.maxes and config.maxes:contiguous()), nil, box), 'b7_sobol_generate')
  return ffi.gc(box[0], B.C.b7_grid_free)
end

return grid
