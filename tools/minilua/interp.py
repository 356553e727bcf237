"""A small Lua 5.1 interpreter (test infrastructure).

No Lua runtime exists in this image, so the LuaJIT glue under lua/bot7_b200/ could never be executed.  This module
runs it: a recursive-descent parser that builds an AST for the complete Lua 5.1 grammar (the token stream is the one
of tools/lua_check.py) and a tree-walking evaluator with the semantics the glue relies on -- lexical scoping with
shared upvalues, multiple assignment / multiple results / varargs, metatables (__index, __newindex, __call, __len,
__eq, __lt, __le, __concat, __unm, arithmetic), string methods through the string metatable, pcall / error with
arbitrary error values, `require` through package.loaded / package.preload / a search path, and the parts of the
base, string, table, math and os libraries that Torch7-style code uses.

Host objects (Torch7 tensors, LuaJIT cdata: see torch7.py, ffi.py) take part through a small protocol:
    lua_index(key)  lua_newindex(key, value)  lua_call(args)  lua_len()  lua_eq(other)  lua_tostring()  lua_type
Python callables are Lua functions: they receive the Lua arguments positionally and return a list of results
(None = no results, any other non-list value = one result).

It is NOT LuaJIT: no coroutines, no goto, no string.dump / loadstring of bytecode, Lua patterns only for the common
cases.  It is the checker's runtime, never the product's.
"""
import math
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lua_check import BINPRI, UNARY_PRI, LuaSyntaxError, tokenize  # noqa: E402


class LuaError(Exception):
    """error(value): carries any Lua value; `traceback` is the chain of Lua call sites."""

    def __init__(self, value, where=None):
        self.value = value
        self.where = where
        super().__init__(tostring(value) if not isinstance(value, str) else value)


class _Break(Exception):
    pass


class _Return(Exception):
    def __init__(self, values):
        self.values = values


# ------------------------------------------------------------------ values

class LuaTable:
    __slots__ = ("hash", "meta", "__weakref__")

    def __init__(self, init=None, meta=None):
        self.hash = dict(init) if init else {}
        self.meta = meta

    def get(self, k):
        if type(k) is float and k.is_integer():
            k = int(k)
        return self.hash.get(k)

    def set(self, k, v):
        if type(k) is float:
            if k != k:
                raise LuaError("table index is NaN")
            if k.is_integer():
                k = int(k)
        if k is None:
            raise LuaError("table index is nil")
        if v is None:
            self.hash.pop(k, None)
        else:
            self.hash[k] = v

    def length(self):
        n = 0
        h = self.hash
        while (n + 1) in h:
            n += 1
        return n

    def __repr__(self):
        return "table: 0x%08x" % (id(self) & 0xFFFFFFFF)


class LuaFunction:
    __slots__ = ("params", "vararg", "body", "scope", "name", "interp", "line")

    def __init__(self, interp, params, vararg, body, scope, name, line):
        self.interp, self.params, self.vararg, self.body, self.scope, self.name, self.line = interp, params, vararg, body, scope, name, line

    def __call__(self, *args):
        return self.interp.call_lua(self, list(args))

    def __repr__(self):
        return "function: %s:%d" % (self.name or "?", self.line)


class Scope:
    __slots__ = ("vars", "parent", "varargs")

    def __init__(self, parent=None):
        self.vars = {}
        self.parent = parent
        self.varargs = None


def lua_type(v):
    if v is None:
        return "nil"
    if v is True or v is False:
        return "boolean"
    if isinstance(v, (int, float)):
        return "number"
    if isinstance(v, str):
        return "string"
    if isinstance(v, LuaTable):
        return "table"
    if isinstance(v, LuaFunction) or callable(v) and not hasattr(v, "lua_type"):
        return "function"
    return getattr(v, "lua_type", "userdata")


def fmt_number(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if v != v:
        return "nan"
    if v in (math.inf, -math.inf):
        return "inf" if v > 0 else "-inf"
    if v.is_integer() and abs(v) < 1e15:
        return "-0" if v == 0 and math.copysign(1.0, v) < 0 else str(int(v))      # printf("%.14g", -0.0)
    return "%.14g" % v


def tostring(v):
    if v is None:
        return "nil"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, (int, float)):
        return fmt_number(v)
    if isinstance(v, str):
        return v
    if isinstance(v, LuaTable):
        mt = v.meta
        if mt is not None and mt.get("__tostring") is not None:
            r = mt.get("__tostring")(v)
            return r[0] if isinstance(r, list) else r
        return repr(v)
    if hasattr(v, "lua_tostring"):
        return v.lua_tostring()
    if isinstance(v, LuaFunction):
        return repr(v)
    if callable(v):
        return "function: builtin: %s" % getattr(v, "__name__", "?")
    return repr(v)


def tonumber(v, base=None):
    if isinstance(v, bool):
        return None
    if isinstance(v, (int, float)):
        return v
    if hasattr(v, "lua_tonumber"):
        return v.lua_tonumber()
    if isinstance(v, str):
        s = v.strip()
        try:
            if base is not None and base != 10:
                return int(s, int(base))
            if s.lower().startswith(("0x", "-0x")):
                return int(s, 16)
            f = float(s)
            return int(f) if f.is_integer() and "." not in s and "e" not in s.lower() and abs(f) < 2 ** 53 else f
        except ValueError:
            return None
    return None


def truthy(v):
    return v is not None and v is not False


def results(r):
    """Normalise what a Python builtin returned into a list of Lua values."""
    if r is None:
        return []
    if isinstance(r, list):
        return r
    if isinstance(r, tuple):
        return list(r)
    return [r]


# ------------------------------------------------------------------ parser -> AST

class Parser:
    """AST nodes are tuples whose first element is the kind; the last element of statements / calls is the line."""

    def __init__(self, text, fname="<lua>"):
        self.fname = fname
        self.toks = tokenize(text, fname)
        self.i = 0

    @property
    def tok(self):
        return self.toks[self.i]

    def err(self, msg):
        raise LuaSyntaxError(f"{self.fname}:{self.tok[2]}: {msg} near {self.tok[1]!r}")

    def check(self, val):
        return self.tok[0] in ("op", "keyword") and self.tok[1] == val

    def accept(self, val):
        if self.check(val):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            self.err(f"{val!r} expected")

    def name(self):
        if self.tok[0] != "name":
            self.err("<name> expected")
        v = self.tok[1]
        self.i += 1
        return v

    def block_end(self):
        return self.tok[0] == "eof" or (self.tok[0] == "keyword" and self.tok[1] in ("end", "else", "elseif", "until"))

    def chunk(self):
        body = self.block()
        if self.tok[0] != "eof":
            self.err("<eof> expected")
        return body

    def block(self):
        stats = []
        while not self.block_end():
            line = self.tok[2]
            if self.accept("return"):
                exps = [] if self.block_end() or self.check(";") else self.explist()
                self.accept(";")
                stats.append(("return", exps, line))
                break
            if self.accept("break"):
                self.accept(";")
                stats.append(("break", line))
                break
            stats.append(self.statement())
            self.accept(";")
        return stats

    def statement(self):
        t = self.tok
        line = t[2]
        if t[0] == "keyword":
            kw = t[1]
            if kw == "if":
                self.i += 1
                clauses = []
                cond = self.exp()
                self.expect("then")
                clauses.append((cond, self.block()))
                orelse = None
                while True:
                    if self.accept("elseif"):
                        cond = self.exp()
                        self.expect("then")
                        clauses.append((cond, self.block()))
                    elif self.accept("else"):
                        orelse = self.block()
                        self.expect("end")
                        break
                    else:
                        self.expect("end")
                        break
                return ("if", clauses, orelse, line)
            if kw == "while":
                self.i += 1
                cond = self.exp()
                self.expect("do")
                body = self.block()
                self.expect("end")
                return ("while", cond, body, line)
            if kw == "do":
                self.i += 1
                body = self.block()
                self.expect("end")
                return ("do", body, line)
            if kw == "for":
                self.i += 1
                n1 = self.name()
                if self.accept("="):
                    a = self.exp()
                    self.expect(",")
                    b = self.exp()
                    c = self.exp() if self.accept(",") else None
                    self.expect("do")
                    body = self.block()
                    self.expect("end")
                    return ("fornum", n1, a, b, c, body, line)
                names = [n1]
                while self.accept(","):
                    names.append(self.name())
                self.expect("in")
                exps = self.explist()
                self.expect("do")
                body = self.block()
                self.expect("end")
                return ("forin", names, exps, body, line)
            if kw == "repeat":
                self.i += 1
                body = self.block()
                self.expect("until")
                cond = self.exp()
                return ("repeat", body, cond, line)
            if kw == "function":
                self.i += 1
                n = self.name()
                target = ("name", n, line)
                full = n
                is_method = False
                while self.check(".") or self.check(":"):
                    colon = self.check(":")
                    self.i += 1
                    key = self.name()
                    full += (":" if colon else ".") + key
                    target = ("index", target, ("str", key), line)
                    if colon:
                        is_method = True
                        break
                fn = self.funcbody(is_method, full, line)
                return ("assign", [target], [fn], line)
            if kw == "local":
                self.i += 1
                if self.accept("function"):
                    n = self.name()
                    return ("localfunc", n, self.funcbody(False, n, line), line)
                names = [self.name()]
                while self.accept(","):
                    names.append(self.name())
                exps = self.explist() if self.accept("=") else []
                return ("local", names, exps, line)
            self.err("unexpected keyword")
        e = self.suffixedexp()
        if self.check("=") or self.check(","):
            targets = [e]
            while self.accept(","):
                targets.append(self.suffixedexp())
            self.expect("=")
            exps = self.explist()
            for tg in targets:
                if tg[0] not in ("name", "index"):
                    self.err("cannot assign to this expression")
            return ("assign", targets, exps, line)
        if e[0] not in ("call", "method"):
            self.err("syntax error (an expression is not a statement)")
        return ("callstat", e, line)

    def funcbody(self, is_method, name, line):
        self.expect("(")
        params = ["self"] if is_method else []
        vararg = False
        if not self.check(")"):
            while True:
                if self.accept("..."):
                    vararg = True
                    break
                params.append(self.name())
                if not self.accept(","):
                    break
        self.expect(")")
        body = self.block()
        self.expect("end")
        return ("func", params, vararg, body, name, line)

    def explist(self):
        exps = [self.exp()]
        while self.accept(","):
            exps.append(self.exp())
        return exps

    def primaryexp(self):
        t = self.tok
        if t[0] == "name":
            self.i += 1
            return ("name", t[1], t[2])
        if self.accept("("):
            e = self.exp()
            self.expect(")")
            return ("paren", e)
        self.err("unexpected symbol")

    def suffixedexp(self):
        e = self.primaryexp()
        while True:
            line = self.tok[2]
            if self.accept("."):
                e = ("index", e, ("str", self.name()), line)
            elif self.accept("["):
                k = self.exp()
                self.expect("]")
                e = ("index", e, k, line)
            elif self.accept(":"):
                m = self.name()
                e = ("method", e, m, self.callargs(), line)
            elif self.check("(") or self.tok[0] == "string" or self.check("{"):
                e = ("call", e, self.callargs(), line)
            else:
                return e

    def callargs(self):
        if self.tok[0] == "string":
            s = self.tok[1]
            self.i += 1
            return [("str", unescape(s))]
        if self.check("{"):
            return [self.table()]
        self.expect("(")
        args = [] if self.check(")") else self.explist()
        self.expect(")")
        return args

    def table(self):
        self.expect("{")
        items = []
        while not self.check("}"):
            if self.accept("["):
                k = self.exp()
                self.expect("]")
                self.expect("=")
                items.append(("kv", k, self.exp()))
            elif self.tok[0] == "name" and self.toks[self.i + 1][0] == "op" and self.toks[self.i + 1][1] == "=":
                k = ("str", self.tok[1])
                self.i += 2
                items.append(("kv", k, self.exp()))
            else:
                items.append(("pos", None, self.exp()))
            if not (self.accept(",") or self.accept(";")):
                break
        self.expect("}")
        return ("table", items)

    def simpleexp(self):
        t = self.tok
        if t[0] == "number":
            self.i += 1
            return ("num", parse_number(t[1]))
        if t[0] == "string":
            self.i += 1
            return ("str", unescape(t[1]))
        if t[0] == "keyword" and t[1] in ("nil", "true", "false"):
            self.i += 1
            return (t[1],)
        if self.accept("..."):
            return ("vararg",)
        if self.check("{"):
            return self.table()
        if self.accept("function"):
            return self.funcbody(False, "anonymous", t[2])
        return self.suffixedexp()

    def exp(self, limit=0):
        t = self.tok
        if (t[0] == "keyword" and t[1] == "not") or (t[0] == "op" and t[1] in ("-", "#")):
            self.i += 1
            e = ("unop", t[1], self.exp(UNARY_PRI), t[2])
        else:
            e = self.simpleexp()
        while True:
            t = self.tok
            op = t[1] if t[0] in ("op", "keyword") else None
            if op not in BINPRI or BINPRI[op][0] <= limit:
                return e
            self.i += 1
            rhs = self.exp(BINPRI[op][1])
            e = (op, e, rhs, t[2]) if op in ("and", "or") else ("binop", op, e, rhs, t[2])


_ESC = {"n": "\n", "t": "\t", "r": "\r", "a": "\a", "b": "\b", "f": "\f", "v": "\v", "\\": "\\", '"': '"', "'": "'", "\n": "\n"}


def unescape(s):
    if "\\" not in s:
        return s
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c != "\\":
            out.append(c)
            i += 1
            continue
        i += 1
        c = s[i]
        if c.isdigit():
            j = i
            while j < len(s) and j < i + 3 and s[j].isdigit():
                j += 1
            out.append(chr(int(s[i:j])))
            i = j
        else:
            out.append(_ESC.get(c, c))
            i += 1
    return "".join(out)


def parse_number(s):
    low = s.lower()
    for suf in ("ull", "ll"):
        if low.endswith(suf):
            return int(low[:-len(suf)], 0)
    if low.startswith("0x"):
        return int(low, 16) if "." not in low and "p" not in low else float.fromhex(low)
    f = float(s)
    return int(f) if re.fullmatch(r"\d+", s) else f


# ------------------------------------------------------------------ evaluator

class Interpreter:
    def __init__(self, search_path=(), stdout=None):
        self.G = LuaTable()
        self.string_meta = LuaTable()
        self.search_path = list(search_path)          # [(module prefix, directory)] e.g. ("bot7_b200", ".../lua/bot7_b200")
        self.stdout = stdout if stdout is not None else sys.stdout
        self.call_depth = 0
        self.chunks = {}
        self.file_stack = []
        from . import stdlib
        stdlib.install(self)

    # ---- running code
    def load(self, text, fname="<lua>"):
        body = Parser(text, fname).chunk()
        return LuaFunction(self, [], True, body, Scope(), fname, 0)

    def run(self, text, fname="<lua>", *args):
        return self.load(text, fname)(*args)

    def run_file(self, path, *args):
        with open(path, encoding="utf-8") as f:
            text = f.read()
        self.file_stack.append(os.path.abspath(path))       # paths.dofile / include resolve against the running file
        try:
            return self.run(text, path, *args)
        finally:
            self.file_stack.pop()

    # ---- calls
    def call(self, f, args, line=None):
        if isinstance(f, LuaFunction):
            return self.call_lua(f, args)
        if isinstance(f, LuaTable):
            h = f.meta.get("__call") if f.meta is not None else None
            if h is None:
                raise LuaError("attempt to call a table value")
            return self.call(h, [f] + args)
        if hasattr(f, "lua_call"):
            return results(f.lua_call(args))
        if callable(f):
            return results(f(*args))
        raise LuaError("attempt to call a %s value" % lua_type(f))

    def call_lua(self, f, args):
        sc = Scope(f.scope)
        np_ = len(f.params)
        for i, p in enumerate(f.params):
            sc.vars[p] = args[i] if i < len(args) else None
        if f.vararg:
            sc.varargs = args[np_:]
        self.call_depth += 1
        if self.call_depth > 180:
            self.call_depth = 0
            raise LuaError("stack overflow")
        try:
            self.exec_block(f.body, sc)
        except _Return as r:
            return r.values
        finally:
            self.call_depth -= 1
        return []

    # ---- metatable-aware primitives
    def index(self, obj, key):
        if isinstance(obj, LuaTable):
            v = obj.get(key)
            if v is not None or obj.meta is None:
                return v
            h = obj.meta.get("__index")
            if h is None:
                return None
            if isinstance(h, LuaTable) or not callable(h):
                return self.index(h, key)
            r = self.call(h, [obj, key])
            return r[0] if r else None
        if isinstance(obj, str):
            return self.index(self.string_meta.get("__index"), key)
        if hasattr(obj, "lua_index"):
            return obj.lua_index(key)
        raise LuaError("attempt to index a %s value (key %s)" % (lua_type(obj), tostring(key)))

    def setindex(self, obj, key, val):
        if isinstance(obj, LuaTable):
            if obj.meta is not None and obj.get(key) is None:
                h = obj.meta.get("__newindex")
                if h is not None:
                    if isinstance(h, LuaTable):
                        return self.setindex(h, key, val)
                    self.call(h, [obj, key, val])
                    return
            obj.set(key, val)
            return
        if hasattr(obj, "lua_newindex"):
            obj.lua_newindex(key, val)
            return
        raise LuaError("attempt to index a %s value (assignment to key %s)" % (lua_type(obj), tostring(key)))

    def metaop(self, name, a, b):
        for v in (a, b):
            if isinstance(v, LuaTable) and v.meta is not None and v.meta.get(name) is not None:
                r = self.call(v.meta.get(name), [a, b])
                return True, (r[0] if r else None)
            if hasattr(v, "lua_arith"):
                r = v.lua_arith(name, a, b)
                if r is not NotImplemented:
                    return True, r
        return False, None

    def arith(self, op, a, b):
        ta, tb = type(a), type(b)
        if (ta is int or ta is float) and (tb is int or tb is float):
            pass
        else:
            na = tonumber(a) if isinstance(a, (str, int, float)) and not isinstance(a, bool) else None
            nb = tonumber(b) if isinstance(b, (str, int, float)) and not isinstance(b, bool) else None
            if na is None or nb is None:
                ok, r = self.metaop({"+": "__add", "-": "__sub", "*": "__mul", "/": "__div", "%": "__mod", "^": "__pow"}[op], a, b)
                if ok:
                    return r
                bad = a if na is None else b
                raise LuaError("attempt to perform arithmetic on a %s value" % lua_type(bad))
            a, b = na, nb
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            try:
                return a / b
            except ZeroDivisionError:
                return math.nan if a == 0 or a != a else math.copysign(math.inf, a) * (math.copysign(1.0, b) if isinstance(b, float) else 1.0)
        if op == "%":
            try:
                return a - math.floor(a / b) * b
            except ZeroDivisionError:
                return math.nan
        if op == "^":
            try:
                return float(a) ** b
            except (OverflowError, ZeroDivisionError):
                return math.inf
        raise LuaError("unknown operator " + op)

    def eq(self, a, b):
        if isinstance(a, float) or isinstance(b, float):
            return (isinstance(a, (int, float)) and isinstance(b, (int, float)) and not isinstance(a, bool)
                    and not isinstance(b, bool) and a == b)          # nan ~= nan
        if a is b:
            return True
        ta, tb = lua_type(a), lua_type(b)
        if ta != tb:
            # LuaJIT: a NULL pointer cdata equals nil
            for x, y in ((a, b), (b, a)):
                if hasattr(x, "lua_eq"):
                    return bool(x.lua_eq(y))
            return False
        if ta in ("number", "string", "boolean"):
            return a == b
        if hasattr(a, "lua_eq"):
            return bool(a.lua_eq(b))
        if isinstance(a, LuaTable) and a.meta is not None and b.meta is not None:
            h = a.meta.get("__eq")
            if h is not None and h is b.meta.get("__eq"):
                r = self.call(h, [a, b])
                return truthy(r[0] if r else None)
        return False

    def less(self, a, b, op="__lt"):
        if isinstance(a, (int, float)) and isinstance(b, (int, float)) and not isinstance(a, bool) and not isinstance(b, bool):
            return a < b if op == "__lt" else a <= b
        if isinstance(a, str) and isinstance(b, str):
            return a < b if op == "__lt" else a <= b
        ok, r = self.metaop(op, a, b)
        if ok:
            return truthy(r)
        raise LuaError("attempt to compare %s with %s" % (lua_type(a), lua_type(b)))

    def concat(self, a, b):
        if isinstance(a, (str, int, float)) and isinstance(b, (str, int, float)) and not isinstance(a, bool) and not isinstance(b, bool):
            return tostring(a) + tostring(b)
        ok, r = self.metaop("__concat", a, b)
        if ok:
            return r
        bad = b if isinstance(a, (str, int, float)) else a
        raise LuaError("attempt to concatenate a %s value" % lua_type(bad))

    def length(self, v):
        if isinstance(v, str):
            return len(v.encode("utf-8", "surrogateescape"))
        if isinstance(v, LuaTable):
            if v.meta is not None and v.meta.get("__len") is not None:        # 5.2 behaviour, harmless
                r = self.call(v.meta.get("__len"), [v])
                return r[0] if r else None
            return v.length()
        if hasattr(v, "lua_len"):
            return v.lua_len()
        raise LuaError("attempt to get length of a %s value" % lua_type(v))

    # ---- statements
    def exec_block(self, stats, scope):
        for st in stats:
            kind = st[0]
            try:
                if kind == "local":
                    vals = self.eval_list(st[2], scope, len(st[1]))
                    for n, v in zip(st[1], vals):
                        scope.vars[n] = v
                elif kind == "assign":
                    targets = st[1]
                    if len(targets) == 1 and len(st[2]) == 1:
                        self.assign(targets[0], self.eval(st[2][0], scope), scope)
                    else:
                        # evaluate the table / key expressions of the targets first, then the values (reference manual 2.4.3)
                        prepared = [self.prepare_target(t, scope) for t in targets]
                        vals = self.eval_list(st[2], scope, len(targets))
                        for p, v in zip(prepared, vals):
                            self.store(p, v, scope)
                elif kind == "callstat":
                    self.eval_multi(st[1], scope)
                elif kind == "if":
                    for cond, body in st[1]:
                        if truthy(self.eval(cond, scope)):
                            self.exec_block(body, Scope(scope))
                            break
                    else:
                        if st[2] is not None:
                            self.exec_block(st[2], Scope(scope))
                elif kind == "fornum":
                    a, b = self.eval(st[2], scope), self.eval(st[3], scope)
                    c = self.eval(st[4], scope) if st[4] is not None else 1
                    a, b, c = tonumber(a), tonumber(b), tonumber(c)
                    if a is None or b is None or c is None:
                        raise LuaError("'for' initial value, limit and step must be numbers")
                    if c == 0:
                        raise LuaError("'for' step is zero")
                    i = a
                    try:
                        while (i <= b) if c > 0 else (i >= b):
                            sc = Scope(scope)
                            sc.vars[st[1]] = i
                            self.exec_block(st[5], sc)
                            i += c
                    except _Break:
                        pass
                elif kind == "forin":
                    vals = self.eval_list(st[2], scope, 3)
                    f, s, ctl = vals[0], vals[1], vals[2]
                    try:
                        while True:
                            rs = self.call(f, [s, ctl])
                            if not rs or rs[0] is None:
                                break
                            ctl = rs[0]
                            sc = Scope(scope)
                            for k, n in enumerate(st[1]):
                                sc.vars[n] = rs[k] if k < len(rs) else None
                            self.exec_block(st[3], sc)
                    except _Break:
                        pass
                elif kind == "while":
                    try:
                        while truthy(self.eval(st[1], scope)):
                            self.exec_block(st[2], Scope(scope))
                    except _Break:
                        pass
                elif kind == "repeat":
                    try:
                        while True:
                            sc = Scope(scope)
                            self.exec_block(st[1], sc)
                            if truthy(self.eval(st[2], sc)):
                                break
                    except _Break:
                        pass
                elif kind == "do":
                    self.exec_block(st[1], Scope(scope))
                elif kind == "localfunc":
                    scope.vars[st[1]] = None
                    scope.vars[st[1]] = self.eval(st[2], scope)
                elif kind == "return":
                    exps = st[1]
                    if len(exps) == 1 and exps[0][0] in ("call", "method"):
                        raise _Return(self.eval_multi(exps[0], scope))          # (tail call)
                    raise _Return(self.eval_list(exps, scope, None))
                elif kind == "break":
                    raise _Break()
                else:
                    raise LuaError("unknown statement " + kind)
            except LuaError as e:
                if e.where is None:
                    e.where = []
                if len(e.where) < 12 and isinstance(st[-1], int):
                    e.where.append(st[-1])
                raise
        return None

    def find_scope(self, name, scope):
        s = scope
        while s is not None:
            if name in s.vars:
                return s
            s = s.parent
        return None

    def prepare_target(self, t, scope):
        if t[0] == "name":
            return ("name", t[1])
        return ("index", self.eval(t[1], scope), self.eval(t[2], scope))

    def store(self, p, v, scope):
        if p[0] == "name":
            s = self.find_scope(p[1], scope)
            if s is not None:
                s.vars[p[1]] = v
            else:
                self.G.set(p[1], v)
        else:
            self.setindex(p[1], p[2], v)

    def assign(self, target, v, scope):
        self.store(self.prepare_target(target, scope), v, scope)

    # ---- expressions
    def eval_list(self, exps, scope, want):
        """Evaluates an expression list; the last expression is expanded if it is a call or `...`.
        want = None: all values; else exactly `want` values (padded with nil / truncated)."""
        vals = []
        n = len(exps)
        for i, e in enumerate(exps):
            if i == n - 1 and e[0] in ("call", "method", "vararg"):
                vals.extend(self.eval_multi(e, scope))
            else:
                vals.append(self.eval(e, scope))
        if want is not None:
            if len(vals) < want:
                vals.extend([None] * (want - len(vals)))
            elif len(vals) > want:
                del vals[want:]
        return vals

    def eval_multi(self, e, scope):
        kind = e[0]
        if kind == "call":
            f = self.eval(e[1], scope)
            args = self.eval_list(e[2], scope, None)
            try:
                return self.call(f, args)
            except LuaError as err:
                if f is None or not (callable(f) or isinstance(f, LuaTable) or hasattr(f, "lua_call")):
                    err.args = ("attempt to call %s (a %s value)" % (describe(e[1]), lua_type(f)),)
                    err.value = err.args[0]
                raise
        if kind == "method":
            obj = self.eval(e[1], scope)
            try:
                f = self.index(obj, e[2])
            except LuaError:
                raise LuaError("attempt to index %s (a %s value) for method '%s'" % (describe(e[1]), lua_type(obj), e[2]))
            if f is None:
                raise LuaError("attempt to call method '%s' (a nil value) on %s" % (e[2], describe(e[1])))
            args = self.eval_list(e[3], scope, None)
            return self.call(f, [obj] + args)
        if kind == "vararg":
            s = scope
            while s is not None and s.varargs is None:
                s = s.parent
            return list(s.varargs) if s is not None else []
        return [self.eval(e, scope)]

    def eval(self, e, scope):
        kind = e[0]
        if kind == "name":
            n = e[1]
            s = scope
            while s is not None:
                if n in s.vars:
                    return s.vars[n]
                s = s.parent
            return self.G.get(n)
        if kind == "num" or kind == "str":
            return e[1]
        if kind == "index":
            obj = self.eval(e[1], scope)
            key = self.eval(e[2], scope)
            if obj is None:
                raise LuaError("attempt to index %s (a nil value) with key '%s'" % (describe(e[1]), tostring(key)))
            return self.index(obj, key)
        if kind == "call" or kind == "method" or kind == "vararg":
            r = self.eval_multi(e, scope)
            return r[0] if r else None
        if kind == "nil":
            return None
        if kind == "true":
            return True
        if kind == "false":
            return False
        if kind == "binop":
            op = e[1]
            a = self.eval(e[2], scope)
            b = self.eval(e[3], scope)
            if op in ("+", "-", "*", "/", "%", "^"):
                return self.arith(op, a, b)
            if op == "==":
                return self.eq(a, b)
            if op == "~=":
                return not self.eq(a, b)
            if op == "<":
                return self.less(a, b)
            if op == "<=":
                return self.less(a, b, "__le")
            if op == ">":
                return self.less(b, a)
            if op == ">=":
                return self.less(b, a, "__le")
            if op == "..":
                return self.concat(a, b)
            raise LuaError("unknown operator " + op)
        if kind == "and":
            a = self.eval(e[1], scope)
            return self.eval(e[2], scope) if truthy(a) else a
        if kind == "or":
            a = self.eval(e[1], scope)
            return a if truthy(a) else self.eval(e[2], scope)
        if kind == "unop":
            v = self.eval(e[2], scope)
            if e[1] == "not":
                return not truthy(v)
            if e[1] == "#":
                return self.length(v)
            if isinstance(v, (int, float)) and not isinstance(v, bool):
                return -v
            if isinstance(v, str) and tonumber(v) is not None:
                return -tonumber(v)
            ok, r = self.metaop("__unm", v, v)
            if ok:
                return r
            raise LuaError("attempt to perform arithmetic on a %s value" % lua_type(v))
        if kind == "func":
            return LuaFunction(self, e[1], e[2], e[3], scope, e[4], e[5])
        if kind == "table":
            t = LuaTable()
            pos = 1
            items = e[1]
            for i, (ik, k, v) in enumerate(items):
                if ik == "kv":
                    t.set(self.eval(k, scope), self.eval(v, scope))
                elif i == len(items) - 1 and v[0] in ("call", "method", "vararg"):
                    for x in self.eval_multi(v, scope):
                        t.set(pos, x)
                        pos += 1
                else:
                    t.set(pos, self.eval(v, scope))
                    pos += 1
            return t
        if kind == "paren":
            return self.eval(e[1], scope)
        raise LuaError("unknown expression " + kind)


def describe(e):
    if e[0] == "name":
        return "'%s'" % e[1]
    if e[0] == "index" and e[2][0] == "str":
        return "field '%s'" % e[2][1]
    if e[0] == "method":
        return "method '%s'" % e[2]
    return "an expression"
