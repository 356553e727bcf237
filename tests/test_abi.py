"""The C-ABI library loads and exports exactly what include/bot7_b200.h declares.  CPU only
(no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bot7_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b7_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from bot7_b200 import _lib
    lib = _lib.load_library()
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bot7_b200.h but not exported"


def test_binding_table_matches_header():
    from bot7_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_lua_cdef_matches_header():
    # the LuaJIT glue binds the same symbols (lua/bot7_b200/ffi.lua carries a cdef generated from the header)
    p = os.path.join(ROOT, "lua", "bot7_b200", "ffi.lua")
    text = open(p).read()
    for n in header_symbols():
        assert re.search(r"\b%s\s*\(" % n, text), f"{n} missing from the ffi.cdef block"


def test_init_fails_loudly_without_gpu():
    from bot7_b200 import _lib
    if _lib.lib().b7_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.B7Error, match="no CUDA device"):
        _lib.Context(0)


def test_missing_library_is_an_error(tmp_path):
    from bot7_b200 import _lib
    with pytest.raises(_lib.B7Error, match="no CPU fallback"):
        _lib.load_library(str(tmp_path / "libbot7_b200.so"))


def test_sobol_directions_host_side(oracle):
    # b7_sobol_directions is pure host code: the device table equals the oracle's
    import ctypes as C
    import numpy as np
    from bot7_b200 import _lib
    for dims in (1, 2, 6, 20, 39):
        out = (C.c_uint32 * (dims * 30))()
        assert _lib.lib().b7_sobol_directions(dims, out) == 0
        assert np.array_equal(np.array(out).reshape(dims, 30), oracle.sobol_direction_integers(dims))
    out = (C.c_uint32 * 30)()
    assert _lib.lib().b7_sobol_directions(40, out) < 0      # grids/sobol.lua:36 dims < 40
