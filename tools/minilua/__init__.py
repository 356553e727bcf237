"""tools/minilua: a small Lua 5.1 interpreter with Torch7 tensor and LuaJIT FFI stand-ins (test infrastructure)."""
from .interp import Interpreter, LuaError, LuaFunction, LuaTable, tostring  # noqa: F401
