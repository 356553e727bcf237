"""Base, string, table, math, os libraries and `require` for tools/minilua (the parts Torch7-style code uses)."""
import math
import os
import re
import time

from .interp import LuaError, LuaTable, lua_type, tonumber, tostring, truthy


# ---- Lua patterns -> Python regular expressions (character classes, anchors, quantifiers, captures; no %b / %f)
_CLASS = {"a": "A-Za-z", "d": "0-9", "l": "a-z", "u": "A-Z", "s": r" \t\n\r\f\v", "w": "A-Za-z0-9", "x": "0-9A-Fa-f",
          "p": r"!-/:-@\[-`{-~", "c": r"\x00-\x1f\x7f"}


def lua_pattern_to_re(pat):
    out, i, n = [], 0, len(pat)
    if pat.startswith("^"):
        out.append("^")
        i = 1
    while i < n:
        c = pat[i]
        if c == "%":
            i += 1
            if i >= n:
                raise LuaError("malformed pattern (ends with '%')")
            d = pat[i]
            if d in _CLASS:
                out.append("[" + _CLASS[d] + "]")
            elif d.lower() in _CLASS and d.isupper():
                out.append("[^" + _CLASS[d.lower()] + "]")
            elif d in "bf":
                raise LuaError("pattern item %%%s is not supported by tools/minilua" % d)
            elif d.isdigit():
                out.append("\\" + d)
            else:
                out.append(re.escape(d))
        elif c == "[":
            j = i + 1
            neg = j < n and pat[j] == "^"
            if neg:
                j += 1
            buf = []
            first = True
            while j < n and (pat[j] != "]" or first):
                first = False
                if pat[j] == "%" and j + 1 < n:
                    d = pat[j + 1]
                    buf.append(_CLASS[d] if d in _CLASS else re.escape(d))
                    j += 2
                else:
                    buf.append("\\" + pat[j] if pat[j] in "\\[]^" else pat[j])
                    j += 1
            out.append("[" + ("^" if neg else "") + "".join(buf) + "]")
            i = j
        elif c == ".":
            out.append("(?s:.)")
        elif c == "-":
            out.append("*?")
        elif c in "*+?":
            out.append(c)
        elif c == "$" and i == n - 1:
            out.append("$")
        elif c in "()":
            out.append(c)
        else:
            out.append(re.escape(c))
        i += 1
    return "".join(out)


_FMT = re.compile(r"%([-+ #0]*)(\d*)(?:\.(\d+))?([cdiouxXeEfgGqs%])")


def lua_format(fmt, *args):
    args = list(args)
    pos = [0]

    def sub(m):
        flags, width, prec, conv = m.group(1), m.group(2), m.group(3), m.group(4)
        if conv == "%":
            return "%"
        if pos[0] >= len(args):
            raise LuaError("bad argument #%d to 'format' (no value)" % (pos[0] + 2))
        v = args[pos[0]]
        pos[0] += 1
        spec = "%" + flags + width + ("." + prec if prec is not None else "")
        if conv in "diouxX":
            nv = tonumber(v)
            if nv is None:
                raise LuaError("bad argument #%d to 'format' (number expected, got %s)" % (pos[0] + 1, lua_type(v)))
            return (spec + ("d" if conv in "iu" else conv)) % int(nv)
        if conv in "eEfgG":
            nv = tonumber(v)
            if nv is None:
                raise LuaError("bad argument #%d to 'format' (number expected, got %s)" % (pos[0] + 1, lua_type(v)))
            return (spec + conv) % float(nv)
        if conv == "c":
            return chr(int(tonumber(v)))
        if conv == "q":
            return '"' + tostring(v).replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n") + '"'
        return (spec + "s") % tostring(v)

    return _FMT.sub(sub, fmt)


def install(I):
    G = I.G
    G.set("_G", G)
    G.set("_VERSION", "Lua 5.1 (tools/minilua)")

    # ------------------------------------------------------------ base
    def l_print(*a):
        I.stdout.write("\t".join(tostring(x) for x in a) + "\n")

    def l_type(*a):
        if not a:
            raise LuaError("bad argument #1 to 'type' (value expected)")
        return lua_type(a[0])

    def l_assert(*a):
        if not a or not truthy(a[0]):
            raise LuaError(a[1] if len(a) > 1 else "assertion failed!")
        return list(a)

    def l_error(v=None, level=1):
        raise LuaError(v)

    def l_pcall(f, *a):
        depth = I.call_depth
        try:
            return [True] + I.call(f, list(a))
        except LuaError as e:
            I.call_depth = depth
            return [False, e.value]
        except RecursionError:
            I.call_depth = depth
            return [False, "stack overflow"]

    def l_xpcall(f, h):
        r = l_pcall(f)
        if r[0]:
            return r
        return [False] + I.call(h, [r[1]])

    def l_select(n, *a):
        if n == "#":
            return len(a)
        n = int(n)
        if n < 0:
            n = len(a) + n + 1
        return list(a[n - 1:])

    def l_next(t, k=None):
        keys = list(t.hash.keys())
        if k is None:
            i = 0
        else:
            if type(k) is float and k.is_integer():
                k = int(k)
            try:
                i = keys.index(k) + 1
            except ValueError:
                raise LuaError("invalid key to 'next'")
        if i >= len(keys):
            return [None]
        return [keys[i], t.hash[keys[i]]]

    def l_pairs(t):
        if not isinstance(t, LuaTable):
            if hasattr(t, "lua_pairs"):
                return t.lua_pairs()
            raise LuaError("bad argument #1 to 'pairs' (table expected, got %s)" % lua_type(t))
        items = list(t.hash.items())       # snapshot: assigning nil to existing fields during traversal is allowed
        state = {"i": 0}

        def it(_t, _k):
            while state["i"] < len(items):
                k, _ = items[state["i"]]
                state["i"] += 1
                v = t.hash.get(k)
                if v is not None:
                    return [k, v]
            return [None]
        return [it, t, None]

    def l_ipairs(t):
        def it(tt, i):
            i = int(i) + 1
            v = I.index(tt, i)
            if v is None:
                return [None]
            return [i, v]
        return [it, t, 0]

    def l_setmetatable(t, mt):
        if not isinstance(t, LuaTable):
            raise LuaError("bad argument #1 to 'setmetatable' (table expected, got %s)" % lua_type(t))
        if t.meta is not None and t.meta.get("__metatable") is not None:
            raise LuaError("cannot change a protected metatable")
        t.meta = mt
        return t

    def l_getmetatable(t=None):
        if isinstance(t, str):
            return I.string_meta
        mt = t.meta if isinstance(t, LuaTable) else getattr(t, "lua_meta", None)
        if mt is not None and mt.get("__metatable") is not None:
            return mt.get("__metatable")
        return [mt]

    def l_unpack(t, i=1, j=None):
        j = I.length(t) if j is None else int(j)
        return [I.index(t, k) for k in range(int(i), j + 1)]

    def l_tonumber(v=None, base=None):
        return [tonumber(v, base)]

    def l_loadstring(s, name=None):
        try:
            return I.load(s, name or "=(loadstring)")
        except Exception as e:  # LuaSyntaxError
            return [None, str(e)]

    for name, f in [("print", l_print), ("type", l_type), ("assert", l_assert), ("error", l_error), ("pcall", l_pcall),
                    ("xpcall", l_xpcall), ("select", l_select), ("next", l_next), ("pairs", l_pairs), ("ipairs", l_ipairs),
                    ("setmetatable", l_setmetatable), ("getmetatable", l_getmetatable), ("unpack", l_unpack),
                    ("tonumber", l_tonumber), ("tostring", lambda v=None: tostring(v)), ("loadstring", l_loadstring),
                    ("rawget", lambda t, k: [t.get(k)]), ("rawset", lambda t, k, v: (t.set(k, v), t)[1]),
                    ("rawequal", lambda a, b: a is b or (lua_type(a) in ("number", "string", "boolean") and lua_type(a) == lua_type(b) and a == b)),
                    ("collectgarbage", lambda *a: 0)]:
        G.set(name, f)

    # ------------------------------------------------------------ string
    S = LuaTable()
    I.string_meta.set("__index", S)

    def s_find(s, pat, init=1, plain=None):
        s, pat = tostring(s), tostring(pat)
        init = int(init)
        if init < 0:
            init = max(len(s) + init + 1, 1)
        if init > len(s) + 1:
            return [None]
        if truthy(plain) or not re.search(r"[\^\$\*\+\?\.\(\)\[\]%\-]", pat):
            k = s.find(pat, init - 1)
            return [None] if k < 0 else [k + 1, k + len(pat)]
        m = re.compile(lua_pattern_to_re(pat)).search(s, init - 1)
        if not m:
            return [None]
        return [m.start() + 1, m.end()] + (list(m.groups()) if m.re.groups else [])

    def s_match(s, pat, init=1):
        s = tostring(s)
        m = re.compile(lua_pattern_to_re(tostring(pat))).search(s, int(init) - 1 if init > 0 else max(len(s) + int(init), 0))
        if not m:
            return [None]
        return list(m.groups()) if m.re.groups else [m.group(0)]

    def s_gmatch(s, pat):
        it = re.compile(lua_pattern_to_re(tostring(pat))).finditer(tostring(s))

        def nxt(*_):
            for m in it:
                return list(m.groups()) if m.re.groups else [m.group(0)]
            return [None]
        return nxt

    def s_gsub(s, pat, repl, n=None):
        s = tostring(s)
        rx = re.compile(lua_pattern_to_re(tostring(pat)))
        count = [0]

        def rep(m):
            count[0] += 1
            whole = m.group(0)
            cap = m.group(1) if m.re.groups else whole
            if isinstance(repl, (str, int, float)):
                r = tostring(repl)
                return re.sub(r"%(.)", lambda mm: (whole if mm.group(1) == "0" else (m.group(int(mm.group(1))) if mm.group(1).isdigit() else mm.group(1))), r)
            if isinstance(repl, LuaTable):
                v = repl.get(cap)
            else:
                rr = I.call(repl, list(m.groups()) if m.re.groups else [whole])
                v = rr[0] if rr else None
            return whole if v is None or v is False else tostring(v)
        out = rx.sub(rep, s, count=0 if n is None else int(n))
        return [out, count[0]]

    def s_sub(s, i=1, j=-1):
        s = tostring(s)
        i, j, n = int(i), int(j), len(s)
        if i < 0:
            i = max(n + i + 1, 1)
        elif i == 0:
            i = 1
        if j < 0:
            j = n + j + 1
        elif j > n:
            j = n
        return s[i - 1:j] if i <= j else ""

    def s_rep(s, n):
        return tostring(s) * max(int(n), 0)

    def s_byte(s, i=1, j=None):
        s = tostring(s)
        j = i if j is None else j
        return [ord(c) for c in s_sub(s, i, j)]

    for name, f in [("find", s_find), ("match", s_match), ("gmatch", s_gmatch), ("gsub", s_gsub), ("sub", s_sub), ("rep", s_rep),
                    ("byte", s_byte), ("char", lambda *a: "".join(chr(int(x)) for x in a)), ("format", lua_format),
                    ("len", lambda s: len(tostring(s))), ("lower", lambda s: tostring(s).lower()), ("upper", lambda s: tostring(s).upper()),
                    ("reverse", lambda s: tostring(s)[::-1])]:
        S.set(name, f)
    G.set("string", S)

    # ------------------------------------------------------------ table
    T = LuaTable()

    def t_insert(t, *a):
        n = I.length(t)
        if len(a) == 1:
            I.setindex(t, n + 1, a[0])
        elif len(a) == 2:
            pos = int(a[0])
            for k in range(n, pos - 1, -1):
                I.setindex(t, k + 1, I.index(t, k))
            I.setindex(t, pos, a[1])
        else:
            raise LuaError("wrong number of arguments to 'insert'")

    def t_remove(t, pos=None):
        n = I.length(t)
        if n == 0:
            return [None]
        pos = n if pos is None else int(pos)
        v = I.index(t, pos)
        for k in range(pos, n):
            I.setindex(t, k, I.index(t, k + 1))
        I.setindex(t, n, None)
        return [v]

    def t_concat(t, sep="", i=1, j=None):
        j = I.length(t) if j is None else int(j)
        parts = []
        for k in range(int(i), j + 1):
            v = I.index(t, k)
            if not isinstance(v, (str, int, float)) or isinstance(v, bool):
                raise LuaError("invalid value (at index %d) in table for 'concat'" % k)
            parts.append(tostring(v))
        return tostring(sep).join(parts)

    def t_sort(t, comp=None):
        import functools
        n = I.length(t)
        vals = [I.index(t, k) for k in range(1, n + 1)]
        if comp is None:
            def cmp(a, b):
                return -1 if I.less(a, b) else (1 if I.less(b, a) else 0)
        else:
            def cmp(a, b):
                if truthy((I.call(comp, [a, b]) or [None])[0]):
                    return -1
                if truthy((I.call(comp, [b, a]) or [None])[0]):
                    return 1
                return 0
        vals.sort(key=functools.cmp_to_key(cmp))
        for k, v in enumerate(vals):
            I.setindex(t, k + 1, v)

    for name, f in [("insert", t_insert), ("remove", t_remove), ("concat", t_concat), ("sort", t_sort),
                    ("getn", lambda t: I.length(t)), ("maxn", lambda t: max([k for k in t.hash if isinstance(k, (int, float))] or [0]))]:
        T.set(name, f)
    G.set("table", T)

    # ------------------------------------------------------------ math
    M = LuaTable()

    def num(f):
        def w(*a):
            try:
                return f(*[float(tonumber(x)) for x in a])
            except (TypeError, ValueError) as e:
                if any(tonumber(x) is None for x in a):
                    raise LuaError("bad argument to math function (number expected)")
                if isinstance(e, ValueError):
                    return math.nan
                raise
            except OverflowError:
                return math.inf
        return w

    def m_floor(x):
        x = tonumber(x)
        return x if x != x or x in (math.inf, -math.inf) else math.floor(x)

    def m_ceil(x):
        x = tonumber(x)
        return x if x != x or x in (math.inf, -math.inf) else math.ceil(x)

    def m_log(x):
        x = float(tonumber(x))
        if x == 0:
            return -math.inf
        return math.log(x) if x > 0 else math.nan

    def m_min(*a):
        r = a[0]
        for x in a[1:]:
            if x < r:
                r = x
        return r

    def m_max(*a):
        r = a[0]
        for x in a[1:]:
            if x > r:
                r = x
        return r

    import random as _random
    rng = _random.Random(0)

    def m_random(m=None, n=None):
        if m is None:
            return rng.random()
        if n is None:
            return rng.randint(1, int(m))
        return rng.randint(int(m), int(n))

    for name, f in [("floor", m_floor), ("ceil", m_ceil), ("sqrt", num(lambda x: math.sqrt(x) if x >= 0 else math.nan)), ("exp", num(math.exp)),
                    ("log", m_log), ("log10", num(math.log10)), ("sin", num(math.sin)), ("cos", num(math.cos)), ("tan", num(math.tan)),
                    ("abs", lambda x: abs(tonumber(x))), ("pow", num(lambda a, b: a ** b)), ("fmod", num(math.fmod)),
                    ("min", m_min), ("max", m_max), ("random", m_random), ("randomseed", lambda s=0: rng.seed(s)),
                    ("atan", num(math.atan)), ("atan2", num(math.atan2)), ("asin", num(math.asin)), ("acos", num(math.acos)),
                    ("modf", lambda x: [float(math.trunc(x)), x - math.trunc(x)]), ("tanh", num(math.tanh))]:
        M.set(name, f)
    M.set("pi", math.pi)
    M.set("huge", math.inf)
    G.set("math", M)

    # ------------------------------------------------------------ os / io
    O = LuaTable()
    O.set("getenv", lambda k: [os.environ.get(tostring(k))])
    O.set("time", lambda *a: int(time.time()))
    O.set("clock", lambda: time.process_time())
    O.set("date", lambda *a: time.strftime("%c"))
    G.set("os", O)
    IO = LuaTable()
    IO.set("write", lambda *a: I.stdout.write("".join(tostring(x) for x in a)))
    G.set("io", IO)

    # ------------------------------------------------------------ package / require
    P = LuaTable()
    loaded, preload = LuaTable(), LuaTable()
    P.set("loaded", loaded)
    P.set("preload", preload)
    P.set("path", "")
    G.set("package", P)
    for lib in ("string", "table", "math", "os", "io"):
        loaded.set(lib, G.get(lib))
    loaded.set("_G", G)

    def l_require(name):
        name = tostring(name)
        v = loaded.get(name)
        if v is not None:
            return v
        loader = preload.get(name)
        if loader is not None:
            r = I.call(loader, [name])
            v = r[0] if r and r[0] is not None else True
            loaded.set(name, v)
            return v
        tried = []
        for prefix, directory in I.search_path:
            if name == prefix:
                cands = [os.path.join(directory, "init.lua")]
            elif name.startswith(prefix + "."):
                rel = name[len(prefix) + 1:].replace(".", os.sep)
                cands = [os.path.join(directory, rel + ".lua"), os.path.join(directory, rel, "init.lua")]
            else:
                continue
            for c in cands:
                tried.append(c)
                if os.path.exists(c):
                    r = I.run_file(c, name)
                    v = r[0] if r and r[0] is not None else (loaded.get(name) if loaded.get(name) is not None else True)
                    loaded.set(name, v)
                    return v
        raise LuaError("module '%s' not found:%s" % (name, "".join("\n\tno file '%s'" % t for t in tried) or "\n\tno search path matches"))

    G.set("require", l_require)

    # ------------------------------------------------------------ paths / include (torch's `paths` package and torch.include)
    PT = LuaTable()

    def here():
        return os.path.dirname(I.file_stack[-1]) if I.file_stack else os.getcwd()

    def p_dofile(name, *a):
        return I.run_file(os.path.join(here(), tostring(name)), *a)

    PT.set("dofile", p_dofile)
    PT.set("thisfile", lambda: I.file_stack[-1] if I.file_stack else None)
    PT.set("dirname", lambda p_: os.path.dirname(tostring(p_)) or ".")
    PT.set("basename", lambda p_: os.path.basename(tostring(p_)))
    PT.set("concat", lambda *a: os.path.join(*[tostring(x) for x in a]))
    PT.set("filep", lambda p_: os.path.isfile(tostring(p_)))
    PT.set("dirp", lambda p_: os.path.isdir(tostring(p_)))
    PT.set("mkdir", lambda p_: (os.makedirs(tostring(p_), exist_ok=True), True)[1])
    G.set("paths", PT)
    loaded.set("paths", PT)
    G.set("include", lambda name: (p_dofile(name), None)[1])
