"""LuaJIT FFI stand-in for tools/minilua, backed by ctypes.

`require('ffi')` returns a table with cdef / load / new / cast / gc / string / copy / fill / sizeof / typeof / istype.
The declarations come from the Lua side's own `ffi.cdef[[ ... ]]` block -- this module PARSES them (opaque struct
typedefs, anonymous enums, function prototypes over int / int64_t / double / pointers), so a call through
`lib.b7_xxx(...)` is checked against exactly what the glue declares to LuaJIT:

* the number of arguments must match the prototype ("wrong number of arguments for function call");
* numbers convert to int / int64_t / double parameters, nil to a NULL pointer, cdata arrays and pointers to pointer
  parameters; a number for a pointer parameter, a table, a tensor (instead of tensor:data()) or a pointer of another
  declared struct type are conversion errors, as in LuaJIT ("cannot convert 'number' to 'double *'");
* cdata arrays are zero-filled and bounds-checked here (LuaJIT would silently corrupt memory);
* ffi.gc finalizers run when the cdata object is garbage collected, or at Runtime.close() in reverse creation order.
"""
import ctypes as C
import re
import weakref

from .interp import LuaError, LuaTable, lua_type, tostring

SCALARS = {"int": C.c_int, "unsigned": C.c_uint, "unsigned int": C.c_uint, "int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64,
           "uint64_t": C.c_uint64, "long long": C.c_longlong, "size_t": C.c_size_t, "double": C.c_double, "float": C.c_float,
           "char": C.c_char, "int8_t": C.c_int8, "uint8_t": C.c_uint8, "bool": C.c_bool, "long": C.c_long, "short": C.c_short}
INTEGRAL = {k for k in SCALARS if k not in ("double", "float")}


def norm_type(t):
    """'const double * const *' -> ('double', 2): base type without qualifiers, pointer depth."""
    depth = t.count("*")
    base = re.sub(r"\bconst\b|\bstruct\b|\*", " ", t)
    base = " ".join(base.split())
    return base, depth


class CPtr:
    """A pointer cdata: `ctype` is the normalised base type, `depth` the number of '*'."""
    lua_type = "cdata"

    def __init__(self, ctype, address, keep=None):
        self.base, self.depth = norm_type(ctype)
        self.address = int(address or 0)
        self.keep = keep
        self.finalizer = None

    def lua_tostring(self):
        return "cdata<%s %s>: 0x%012x" % (self.base, "*" * self.depth, self.address)

    def lua_eq(self, other):
        if other is None:
            return self.address == 0
        return isinstance(other, (CPtr, CArray)) and other.addr() == self.address

    def addr(self):
        return self.address

    def lua_index(self, key):
        if isinstance(key, str):
            raise LuaError("'%s %s' has no member named '%s' (opaque type)" % (self.base, "*" * self.depth, key))
        if self.address == 0:
            raise LuaError("attempt to dereference a NULL pointer")
        if self.depth == 1 and self.base in SCALARS:
            ct = SCALARS[self.base]
            v = ct.from_address(self.address + int(key) * C.sizeof(ct)).value
            return v if not isinstance(v, bytes) else v[0]
        if self.depth >= 2:
            v = C.c_void_p.from_address(self.address + int(key) * C.sizeof(C.c_void_p)).value
            return CPtr(self.base + "*" * (self.depth - 1), v or 0)
        raise LuaError("cannot dereference a pointer to the opaque type '%s'" % self.base)

    def lua_newindex(self, key, val):
        if self.depth == 1 and self.base in SCALARS:
            ct = SCALARS[self.base]
            ct.from_address(self.address + int(key) * C.sizeof(ct)).value = val
            return
        raise LuaError("cannot assign through this pointer")


class CArray:
    lua_type = "cdata"

    def __init__(self, base, depth, n, vla):
        self.base, self.depth, self.n, self.vla = base, depth, n, vla
        self.ct = C.c_void_p if depth > 0 else SCALARS[base]
        self.buf = (self.ct * max(n, 1))()
        self.keep = [None] * max(n, 1)
        self.finalizer = None

    def addr(self):
        return C.addressof(self.buf)

    def lua_tostring(self):
        return "cdata<%s %s[%d]>: 0x%012x" % (self.base, "*" * self.depth, self.n, self.addr())

    def lua_eq(self, other):
        return other is self

    def _k(self, key):
        if isinstance(key, bool) or not isinstance(key, (int, float)) or key != int(key):
            raise LuaError("cdata array index must be an integer, got %s" % tostring(key))
        k = int(key)
        if not 0 <= k < self.n:
            raise LuaError("cdata array index %d out of bounds [0, %d) -- LuaJIT would not check this" % (k, self.n))
        return k

    def lua_index(self, key):
        k = self._k(key)
        if self.depth > 0:
            return CPtr(self.base + "*" * self.depth, self.buf[k] or 0, keep=self.keep[k])
        v = self.buf[k]
        return v

    def lua_newindex(self, key, val):
        k = self._k(key)
        if self.depth > 0:
            if val is None:
                self.buf[k], self.keep[k] = None, None
            elif isinstance(val, (CPtr, CArray)):
                check_pointer_compat(val, self.base, self.depth, "array element")
                self.buf[k], self.keep[k] = val.addr(), val
            else:
                raise LuaError("cannot convert '%s' to '%s %s'" % (lua_type(val), self.base, "*" * self.depth))
            return
        if isinstance(val, bool) or not isinstance(val, (int, float)):
            raise LuaError("cannot convert '%s' to '%s'" % (lua_type(val), self.base))
        self.buf[k] = int(val) if self.base in INTEGRAL else float(val)


def check_pointer_compat(val, base, depth, what):
    vb = val.base
    vd = val.depth + (1 if isinstance(val, CArray) else 0)
    if base == "void" or vb == "void":
        return
    if vb != base or vd != depth:
        raise LuaError("cannot convert '%s %s' to '%s %s' (%s)" % (vb, "*" * vd, base, "*" * depth, what))


class CFunc:
    lua_type = "cdata"

    def __init__(self, name, fn, ret, params, log=None):
        self.name, self.fn, self.ret, self.params, self.log = name, fn, ret, params, log
        rb, rd = norm_type(ret)
        if rd > 0:
            fn.restype = C.c_char_p if (rb, rd) == ("char", 1) else C.c_void_p
        elif rb == "void":
            fn.restype = None
        else:
            fn.restype = SCALARS[rb]
        at = []
        for _, (b, d) in params:
            at.append(C.c_void_p if d > 0 else SCALARS[b])
        fn.argtypes = at

    def lua_tostring(self):
        return "cdata<%s ()>: %s" % (self.ret, self.name)

    def lua_call(self, args):
        if len(args) != len(self.params):
            raise LuaError("wrong number of arguments for function call (%s takes %d, got %d)" % (self.name, len(self.params), len(args)))
        conv = []
        for k, (a, (pname, (b, d))) in enumerate(zip(args, self.params)):
            what = "argument #%d '%s' of %s" % (k + 1, pname, self.name)
            if d > 0:
                if a is None:
                    conv.append(None)
                elif isinstance(a, (CPtr, CArray)):
                    check_pointer_compat(a, b, d, what)
                    conv.append(a.addr() or None)
                elif isinstance(a, str) and (b, d) == ("char", 1):
                    conv.append(C.cast(C.c_char_p(a.encode()), C.c_void_p))
                else:
                    raise LuaError("cannot convert '%s' to '%s %s' (%s)" % (lua_type(a), b, "*" * d, what))
            else:
                if isinstance(a, bool):
                    a = int(a)
                if not isinstance(a, (int, float)):
                    raise LuaError("cannot convert '%s' to '%s' (%s)" % (lua_type(a), b, what))
                if b in INTEGRAL:
                    if a != a or a in (float("inf"), float("-inf")):
                        raise LuaError("cannot convert a non-finite number to '%s' (%s)" % (b, what))
                    conv.append(int(a))            # LuaJIT truncates
                else:
                    conv.append(float(a))
        if self.log is not None:
            self.log.append(self.name)
        r = self.fn(*conv)
        rb, rd = norm_type(self.ret)
        if rb == "void" and rd == 0:
            return []
        if (rb, rd) == ("char", 1):
            return [CStr(r)]
        if rd > 0:
            return [CPtr(self.ret, r or 0)]
        return [r]


class CStr:
    lua_type = "cdata"

    def __init__(self, b):
        self.b = b

    def lua_eq(self, other):
        return other is None and self.b is None


class CLib:
    lua_type = "userdata"

    def __init__(self, dll, decls, name, log=None):
        self.dll, self.decls, self.name, self.cache, self.log = dll, decls, name, {}, log

    def lua_index(self, key):
        if key in self.decls.enums:
            return self.decls.enums[key]
        if key in self.cache:
            return self.cache[key]
        if key not in self.decls.funcs:
            raise LuaError("missing declaration for symbol '%s'" % tostring(key))
        try:
            fn = getattr(self.dll, key)
        except AttributeError:
            raise LuaError("cannot resolve symbol '%s': undefined symbol in %s" % (key, self.name))
        ret, params = self.decls.funcs[key]
        f = CFunc(key, fn, ret, params, self.log)
        self.cache[key] = f
        return f


class Decls:
    def __init__(self):
        self.structs, self.enums, self.funcs = set(), {}, {}

    def parse(self, text):
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
        text = re.sub(r"//[^\n]*", " ", text)
        for stmt in text.split(";"):
            s = " ".join(stmt.split())
            if not s:
                continue
            m = re.fullmatch(r"typedef struct (\w+) (\w+)", s)
            if m:
                self.structs.add(m.group(2))
                continue
            m = re.fullmatch(r"enum(?: \w+)? ?\{(.*)\}", s)
            if m:
                nxt = 0
                for item in m.group(1).split(","):
                    item = item.strip()
                    if not item:
                        continue
                    if "=" in item:
                        k, v = item.split("=")
                        nxt = int(v.strip(), 0)
                        k = k.strip()
                    else:
                        k = item
                    self.enums[k] = nxt
                    nxt += 1
                continue
            m = re.fullmatch(r"(.+?)\b(\w+) ?\((.*)\)", s)
            if m:
                ret, name, plist = m.group(1).strip(), m.group(2), m.group(3).strip()
                params = []
                if plist and plist != "void":
                    for k, p in enumerate(plist.split(",")):
                        p = p.strip()
                        pm = re.fullmatch(r"(.+?[\s\*])(\w+)", p)
                        if pm and pm.group(2) not in SCALARS and pm.group(2) not in self.structs and pm.group(2) != "const":
                            ptype, pname = pm.group(1).strip(), pm.group(2)
                        else:
                            ptype, pname = p, "arg%d" % (k + 1)
                        b, d = norm_type(ptype)
                        if b not in SCALARS and b not in self.structs and b != "void":
                            raise LuaError("ffi.cdef: unknown type '%s' in the declaration of %s" % (b, name))
                        params.append((pname, (b, d)))
                rb, _ = norm_type(ret)
                if rb not in SCALARS and rb not in self.structs and rb != "void":
                    raise LuaError("ffi.cdef: unknown return type '%s' of %s" % (rb, name))
                self.funcs[name] = (ret, params)
                continue
            raise LuaError("ffi.cdef: cannot parse declaration '%s'" % s)


class Runtime:
    """One per interpreter: owns the declarations, the finalizer list and the library loader."""

    def __init__(self, interp, lib_resolver):
        self.I = interp
        self.decls = Decls()
        self.lib_resolver = lib_resolver       # name -> ctypes.CDLL (raises OSError)
        self.finalizers = []                   # weakref.finalize objects, creation order
        self.calls = []                        # names of the C functions called (for the tests)

    def parse_ctype(self, ct):
        """'b7_gp*[1]' / 'int[?]' / 'const double*[?]' / 'double[4]' / 'int64_t' -> (base, depth, n or None, vla)"""
        m = re.fullmatch(r"\s*(.+?)\s*(?:\[\s*(\?|\d+)\s*\])?\s*", ct)
        if not m:
            raise LuaError("ffi: cannot parse ctype '%s'" % ct)
        b, d = norm_type(m.group(1))
        if b not in SCALARS and b not in self.decls.structs and b != "void":
            raise LuaError("ffi: undeclared type '%s' in '%s'" % (b, ct))
        if m.group(2) is None:
            return b, d, None, False
        if m.group(2) == "?":
            return b, d, None, True
        return b, d, int(m.group(2)), False

    def new(self, ct, *init):
        b, d, n, vla = self.parse_ctype(tostring(ct))
        init = list(init)
        if vla:
            if not init or isinstance(init[0], bool) or not isinstance(init[0], (int, float)):
                raise LuaError("ffi.new('%s'): the size of the VLA is missing" % ct)
            n = int(init.pop(0))
            if n < 0:
                raise LuaError("ffi.new: negative array size")
        if n is None:
            n, scalar = 1, True            # a scalar or a single pointer: boxed as an array of one
        else:
            scalar = False
        if d == 0 and b not in SCALARS:
            raise LuaError("ffi.new: cannot instantiate the opaque type '%s'" % b)
        arr = CArray(b, d, n, vla)
        if init:
            if len(init) == 1 and isinstance(init[0], LuaTable):
                vals = [init[0].get(k) for k in range(1, init[0].length() + 1)]
            else:
                vals = init
            if len(vals) > n:
                raise LuaError("ffi.new: too many initializers (%d for %d elements)" % (len(vals), n))
            for k, v in enumerate(vals):
                arr.lua_newindex(k, v)
            if len(vals) == 1 and not isinstance(init[0], LuaTable) and n > 1:     # a single initializer fills the array
                for k in range(1, n):
                    arr.lua_newindex(k, vals[0])
        del scalar
        return arr

    def gc(self, cdata, fin):
        if not isinstance(cdata, (CPtr, CArray)):
            raise LuaError("ffi.gc: cdata expected, got %s" % lua_type(cdata))
        if fin is None:
            if cdata.finalizer is not None:
                cdata.finalizer.detach()
                cdata.finalizer = None
            return cdata
        if isinstance(cdata, CPtr):
            ghost = CPtr(cdata.base + "*" * cdata.depth, cdata.address)      # what the finalizer receives
        else:
            ghost = cdata
        interp = self.I

        def run():
            interp.call(fin, [ghost])
        f = weakref.finalize(cdata, run) if isinstance(cdata, CPtr) else None
        if f is not None:
            f.atexit = False
            cdata.finalizer = f
            self.finalizers.append(f)
        return cdata

    def close(self):
        import gc
        gc.collect()
        while self.finalizers:
            f = self.finalizers.pop()
            if f.alive:
                f()

    def install(self):
        I, F = self.I, LuaTable()

        def f_cdef(text):
            self.decls.parse(tostring(text))

        def f_load(name, _global=None):
            try:
                dll = self.lib_resolver(tostring(name))
            except OSError as e:
                raise LuaError("cannot load library '%s': %s" % (tostring(name), e))
            return CLib(dll, self.decls, tostring(name), self.calls)

        def f_string(p, n=None):
            if isinstance(p, CStr):
                return p.b.decode() if p.b is not None else None
            if isinstance(p, (CPtr, CArray)):
                if p.addr() == 0:
                    raise LuaError("ffi.string: NULL pointer")
                return (C.string_at(p.addr(), int(n)) if n is not None else C.string_at(p.addr())).decode(errors="replace")
            raise LuaError("ffi.string: cdata expected, got %s" % lua_type(p))

        def f_sizeof(ct, n=None):
            if isinstance(ct, CArray):
                return C.sizeof(ct.ct) * ct.n
            if isinstance(ct, CPtr):
                return C.sizeof(C.c_void_p)
            b, d, cnt, vla = self.parse_ctype(tostring(ct))
            one = C.sizeof(C.c_void_p) if d > 0 else C.sizeof(SCALARS[b])
            if vla:
                cnt = int(n)
            return one * (cnt if cnt is not None else 1)

        def f_copy(dst, src, n=None):
            if isinstance(src, str):
                raw = src.encode() + b"\0"
                C.memmove(dst.addr(), raw, len(raw) if n is None else int(n))
                return
            if n is None:
                raise LuaError("ffi.copy: length expected")
            n = int(n)
            for what, x in (("destination", dst), ("source", src)):
                if not isinstance(x, (CPtr, CArray)):
                    raise LuaError("ffi.copy: %s must be cdata, got %s" % (what, lua_type(x)))
                if isinstance(x, CArray) and n > C.sizeof(x.ct) * x.n:
                    raise LuaError("ffi.copy: %d bytes exceed the %s array (%d bytes)" % (n, what, C.sizeof(x.ct) * x.n))
            C.memmove(dst.addr(), src.addr(), n)

        def f_fill(dst, n, c=0):
            C.memset(dst.addr(), int(c), int(n))

        def f_cast(ct, v):
            b, d, cnt, vla = self.parse_ctype(tostring(ct))
            if d == 0:
                return int(v) if b in INTEGRAL else float(v)
            if v is None:
                return CPtr(b + "*" * d, 0)
            if isinstance(v, (CPtr, CArray)):
                return CPtr(b + "*" * d, v.addr(), keep=v)
            if isinstance(v, (int, float)):
                return CPtr(b + "*" * d, int(v))
            raise LuaError("ffi.cast: cannot convert '%s'" % lua_type(v))

        def f_istype(ct, v):
            if not isinstance(v, (CPtr, CArray)):
                return False
            b, d, cnt, vla = self.parse_ctype(tostring(ct))
            return v.base == b

        for name, f in [("cdef", f_cdef), ("load", f_load), ("new", self.new), ("gc", self.gc), ("string", f_string), ("sizeof", f_sizeof),
                        ("copy", f_copy), ("fill", f_fill), ("cast", f_cast), ("istype", f_istype), ("typeof", lambda ct: ct)]:
            F.set(name, f)
        F.set("os", "Linux")
        F.set("arch", "x64")
        I.G.get("package").get("loaded").set("ffi", F)
        return F
