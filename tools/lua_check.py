"""A Lua 5.1 / LuaJIT syntax and scope checker (no Lua runtime exists in this image).

The LuaJIT glue under lua/bot7_b200/ cannot be executed here; this parses it with the complete Lua 5.1 grammar
(Reference Manual section 8, plus LuaJIT's hexadecimal / LL / ULL number literals) and resolves every name:

* a syntax error is reported with file:line;
* every name that is read is either a local in scope (local, function parameter, loop variable, implicit `self` of a
  `function a:b()` definition, `...` only inside a vararg function) or a GLOBAL read, which the caller compares with
  the set of globals a Torch7 / LuaJIT process provides -- a misspelt local shows up as an unknown global;
* global assignments and the method / field names called on each object are collected for the tests.

usage: python tools/lua_check.py file.lua ...      (exit 1 on a syntax error or an unknown global)
"""
import re
import sys

KEYWORDS = {"and", "break", "do", "else", "elseif", "end", "false", "for", "function", "if", "in", "local", "nil", "not", "or",
            "repeat", "return", "then", "true", "until", "while"}

# what a `th` / LuaJIT process with torch loaded provides (Lua 5.1 base library, LuaJIT's, Torch7's)
KNOWN_GLOBALS = {"_G", "_VERSION", "assert", "collectgarbage", "dofile", "error", "getfenv", "getmetatable", "ipairs", "load",
                 "loadfile", "loadstring", "module", "next", "pairs", "pcall", "print", "rawequal", "rawget", "rawset", "require",
                 "select", "setfenv", "setmetatable", "tonumber", "tostring", "type", "unpack", "xpcall", "coroutine", "debug", "io",
                 "math", "os", "package", "string", "table", "bit", "jit", "arg", "include",
                 "torch", "nn", "paths", "xlua", "bot7", "optim", "gnuplot", "sys", "cutorch", "cunn", "gpTorch7", "gp"}

TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lcomment>--\[(?P<lc_eq>=*)\[)
  | (?P<comment>--[^\n]*)
  | (?P<lstring>\[(?P<ls_eq>=*)\[)
  | (?P<number>0[xX][0-9a-fA-F]+(?:\.[0-9a-fA-F]*)?(?:[pP][+-]?\d+)?(?:ULL|LL|ull|ll|i)?
              |(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?(?:ULL|LL|ull|ll|i)?)
  | (?P<name>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<string>"(?:\\.|\\\n|[^"\\\n])*"|'(?:\\.|\\\n|[^'\\\n])*')
  | (?P<op>\.\.\.|\.\.|==|~=|<=|>=|[-+*/%^\#<>=(){}\[\];:,.])
""", re.X | re.S)


class LuaSyntaxError(Exception):
    pass


def tokenize(text, fname="<lua>"):
    toks, pos, line = [], 0, 1
    if text.startswith("#"):                       # shebang line
        pos = text.index("\n") if "\n" in text else len(text)
    while pos < len(text):
        m = TOKEN.match(text, pos)
        if not m:
            raise LuaSyntaxError(f"{fname}:{line}: unexpected character {text[pos]!r}")
        kind = m.lastgroup if m.lastgroup not in ("lc_eq", "ls_eq") else None
        if m.group("lcomment") is not None or m.group("lstring") is not None:
            is_comment = m.group("lcomment") is not None
            eq = m.group("lc_eq") if is_comment else m.group("ls_eq")
            close = "]" + eq + "]"
            end = text.find(close, m.end())
            if end < 0:
                raise LuaSyntaxError(f"{fname}:{line}: unfinished long {'comment' if is_comment else 'string'}")
            body = text[m.end():end]
            if not is_comment:
                toks.append(("string", body, line))
            line += text.count("\n", pos, end + len(close))
            pos = end + len(close)
            continue
        val = m.group(0)
        if m.group("ws") is not None or m.group("comment") is not None:
            pass
        elif m.group("number") is not None:
            toks.append(("number", val, line))
        elif m.group("name") is not None:
            toks.append(("keyword" if val in KEYWORDS else "name", val, line))
        elif m.group("string") is not None:
            toks.append(("string", val[1:-1], line))
        else:
            toks.append(("op", val, line))
        line += val.count("\n")
        pos = m.end()
        del kind
    toks.append(("eof", "<eof>", line))
    return toks


BINPRI = {"or": (1, 1), "and": (2, 2), "<": (3, 3), ">": (3, 3), "<=": (3, 3), ">=": (3, 3), "~=": (3, 3), "==": (3, 3),
          "..": (5, 4), "+": (6, 6), "-": (6, 6), "*": (7, 7), "/": (7, 7), "%": (7, 7), "^": (10, 9)}   # (left, right) priorities
UNARY_PRI = 8


class Scope:
    def __init__(self, parent=None, function=False, vararg=False):
        self.parent, self.names, self.function, self.vararg = parent, set(), function, vararg

    def lookup(self, name):
        s = self
        while s:
            if name in s.names:
                return True
            s = s.parent
        return False

    def in_vararg_function(self):
        s = self
        while s and not s.function:
            s = s.parent
        return bool(s and s.vararg)


class Report:
    def __init__(self, fname):
        self.fname = fname
        self.global_reads = {}        # name -> first line
        self.global_writes = {}       # name -> first line
        self.method_calls = []        # (object expression text, method, line)   obj:method(...)
        self.field_calls = []         # (dotted path, line, number of argument expressions)   a.b.c(...)
        self.functions = []           # (dotted name incl. ':' for methods, line)
        self.locals_declared = 0


class Parser:
    def __init__(self, text, fname="<lua>"):
        self.fname = fname
        self.toks = tokenize(text, fname)
        self.i = 0
        self.rep = Report(fname)
        self.scope = Scope(function=True, vararg=True)      # the main chunk is a vararg function
        self.loop_depth = 0

    # ---- token helpers
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

    def expect(self, val, what=None):
        if not self.accept(val):
            self.err(f"{what or repr(val)} expected")

    def name(self):
        if self.tok[0] != "name":
            self.err("<name> expected")
        v = self.tok[1]
        self.i += 1
        return v

    # ---- scopes
    def push(self, function=False, vararg=False):
        self.scope = Scope(self.scope, function, vararg)

    def pop(self):
        self.scope = self.scope.parent

    def declare(self, name):
        self.scope.names.add(name)
        self.rep.locals_declared += 1

    def read(self, name, line):
        if not self.scope.lookup(name):
            self.rep.global_reads.setdefault(name, line)

    # ---- grammar
    def chunk(self):
        self.block()
        if self.tok[0] != "eof":
            self.err("<eof> expected")
        return self.rep

    def block_end(self):
        return self.tok[0] == "eof" or (self.tok[0] == "keyword" and self.tok[1] in ("end", "else", "elseif", "until"))

    def block(self, scoped=True):
        if scoped:
            self.push()
        while not self.block_end():
            if self.check("return"):
                self.i += 1
                if not self.block_end() and not self.check(";"):
                    self.explist()
                self.accept(";")
                if not self.block_end():
                    self.err("'return' must be the last statement of its block; <eof> or 'end'")
                break
            if self.check("break"):
                if self.loop_depth == 0:
                    self.err("'break' outside a loop")
                self.i += 1
                self.accept(";")
                if not self.block_end():
                    self.err("'break' must be the last statement of its block; 'end'")
                break
            self.statement()
            self.accept(";")
        if scoped:
            self.pop()

    def statement(self):
        t = self.tok
        if t[0] == "keyword":
            kw = t[1]
            if kw == "if":
                self.i += 1
                self.exp()
                self.expect("then")
                self.block()
                while self.check("elseif"):
                    self.i += 1
                    self.exp()
                    self.expect("then")
                    self.block()
                if self.accept("else"):
                    self.block()
                self.expect("end", "'end' (to close 'if' at line %d)" % t[2])
                return
            if kw == "while":
                self.i += 1
                self.exp()
                self.expect("do")
                self.loop_depth += 1
                self.block()
                self.loop_depth -= 1
                self.expect("end", "'end' (to close 'while' at line %d)" % t[2])
                return
            if kw == "do":
                self.i += 1
                self.block()
                self.expect("end", "'end' (to close 'do' at line %d)" % t[2])
                return
            if kw == "for":
                self.i += 1
                n1 = self.name()
                self.push()
                if self.accept("="):
                    self.exp()
                    self.expect(",")
                    self.exp()
                    if self.accept(","):
                        self.exp()
                    names = [n1]
                else:
                    names = [n1]
                    while self.accept(","):
                        names.append(self.name())
                    self.expect("in", "'=' or 'in'")
                    self.explist()
                self.expect("do")
                for n in names:
                    self.declare(n)
                self.loop_depth += 1
                self.block()
                self.loop_depth -= 1
                self.pop()
                self.expect("end", "'end' (to close 'for' at line %d)" % t[2])
                return
            if kw == "repeat":
                self.i += 1
                self.push()
                self.loop_depth += 1
                self.block(scoped=False)            # the condition sees the body's locals
                self.loop_depth -= 1
                self.expect("until", "'until' (to close 'repeat' at line %d)" % t[2])
                self.exp()
                self.pop()
                return
            if kw == "function":
                self.i += 1
                line = self.tok[2]
                n = self.name()
                self.read(n, line) if not self.scope.lookup(n) and (self.check(".") or self.check(":")) else None
                path, is_method = n, False
                first_is_global = not self.scope.lookup(n)
                while self.check("."):
                    self.i += 1
                    path += "." + self.name()
                if self.accept(":"):
                    path += ":" + self.name()
                    is_method = True
                if path == n and first_is_global:
                    self.rep.global_writes.setdefault(n, line)
                self.rep.functions.append((path, line))
                self.funcbody(is_method, line)
                return
            if kw == "local":
                self.i += 1
                if self.accept("function"):
                    line = self.tok[2]
                    n = self.name()
                    self.declare(n)                 # visible inside its own body (recursion)
                    self.rep.functions.append((n, line))
                    self.funcbody(False, line)
                    return
                names = [self.name()]
                while self.accept(","):
                    names.append(self.name())
                if self.accept("="):
                    self.explist()
                for n in names:                     # in scope only after the statement
                    self.declare(n)
                return
            self.err("unexpected keyword")
        # exprstat: assignment or call
        line = t[2]
        kind, info = self.suffixedexp(assign_target=True)
        if self.check("=") or self.check(","):
            targets = [(kind, info)]
            while self.accept(","):
                targets.append(self.suffixedexp(assign_target=True))
            self.expect("=")
            self.explist()
            for k, inf in targets:
                if k == "global":
                    self.rep.global_writes.setdefault(inf, line)
                elif k == "call" or k == "other":
                    self.err("cannot assign to this expression")
        else:
            if kind == "global":
                self.rep.global_reads.setdefault(info, line)
            if kind != "call":
                self.err("syntax error (an expression is not a statement)")

    def funcbody(self, is_method, line):
        self.expect("(")
        self.push(function=True)
        saved_loop, self.loop_depth = self.loop_depth, 0
        if is_method:
            self.declare("self")
        if not self.check(")"):
            while True:
                if self.accept("..."):
                    self.scope.vararg = True
                    self.declare("arg")            # Lua 5.1 compatibility vararg table
                    break
                self.declare(self.name())
                if not self.accept(","):
                    break
        self.expect(")")
        self.block(scoped=False)
        self.expect("end", "'end' (to close 'function' at line %d)" % line)
        self.loop_depth = saved_loop
        self.pop()

    def explist(self):
        n = 1
        self.exp()
        while self.accept(","):
            self.exp()
            n += 1
        return n

    def primaryexp(self, assign_target):
        t = self.tok
        if t[0] == "name":
            self.i += 1
            if self.scope.lookup(t[1]):
                return "local", t[1], t[1]
            return "global", t[1], t[1]
        if self.accept("("):
            self.exp()
            self.expect(")")
            return "other", None, "(...)"
        self.err("unexpected symbol")

    def suffixedexp(self, assign_target=False):
        """Returns (kind, info): kind in local / global (a bare name; the caller decides read or write), index, call, other."""
        line = self.tok[2]
        kind, info, path = self.primaryexp(assign_target)
        bare = kind in ("local", "global")
        while True:
            if self.check("."):
                if bare and kind == "global":
                    self.rep.global_reads.setdefault(info, line)
                self.i += 1
                path = (path + "." if path else "") + self.name() if path is not None else None
                kind, bare = "index", False
            elif self.check("["):
                if bare and kind == "global":
                    self.rep.global_reads.setdefault(info, line)
                self.i += 1
                self.exp()
                self.expect("]")
                path = path + "[]" if path is not None else None
                kind, bare = "index", False
            elif self.check(":"):
                if bare and kind == "global":
                    self.rep.global_reads.setdefault(info, line)
                self.i += 1
                m = self.name()
                self.rep.method_calls.append((path, m, self.tok[2]))
                self.callargs()
                path = (path or "") + ":" + m + "()"
                kind, bare = "call", False
            elif self.check("(") or self.tok[0] == "string" or self.check("{"):
                if bare and kind == "global":
                    self.rep.global_reads.setdefault(info, line)
                call_line = self.tok[2]
                n_args = self.callargs()
                if path is not None:
                    self.rep.field_calls.append((path, call_line, n_args))
                path = (path or "") + "()"
                kind, bare = "call", False
            else:
                break
        if bare:
            return kind, info
        return kind, path

    def callargs(self):
        """Returns the number of argument expressions (a trailing call or `...` may expand to more at run time)."""
        if self.tok[0] == "string":
            self.i += 1
            return 1
        if self.check("{"):
            self.table()
            return 1
        line = self.tok[2]
        n = 0
        self.expect("(", "function arguments")
        if not self.check(")"):
            n = self.explist()
        self.expect(")", "')' (to close '(' at line %d)" % line)
        return n

    def table(self):
        line = self.tok[2]
        self.expect("{")
        while not self.check("}"):
            if self.check("["):
                self.i += 1
                self.exp()
                self.expect("]")
                self.expect("=")
                self.exp()
            elif self.tok[0] == "name" and self.toks[self.i + 1][0] == "op" and self.toks[self.i + 1][1] == "=":
                self.i += 2
                self.exp()
            else:
                self.exp()
            if not (self.accept(",") or self.accept(";")):
                break
        self.expect("}", "'}' (to close '{' at line %d)" % line)

    def simpleexp(self):
        t = self.tok
        if t[0] in ("number", "string"):
            self.i += 1
        elif t[0] == "keyword" and t[1] in ("nil", "true", "false"):
            self.i += 1
        elif self.check("..."):
            if not self.scope.in_vararg_function():
                self.err("cannot use '...' outside a vararg function")
            self.i += 1
        elif self.check("{"):
            self.table()
        elif self.check("function"):
            self.i += 1
            self.funcbody(False, t[2])
        else:
            kind, info = self.suffixedexp()
            if kind == "global":
                self.rep.global_reads.setdefault(info, t[2])

    def exp(self, limit=0):
        t = self.tok
        if (t[0] == "keyword" and t[1] == "not") or (t[0] == "op" and t[1] in ("-", "#")):
            self.i += 1
            self.exp(UNARY_PRI)
        else:
            self.simpleexp()
        while True:
            t = self.tok
            op = t[1] if t[0] in ("op", "keyword") else None
            if op not in BINPRI or BINPRI[op][0] <= limit:
                break
            self.i += 1
            self.exp(BINPRI[op][1])


def check_text(text, fname="<lua>"):
    return Parser(text, fname).chunk()


def check_file(path):
    with open(path, encoding="utf-8", errors="replace") as f:
        return check_text(f.read(), path)


def unknown_globals(rep, extra_known=()):
    """Global names that are read but neither provided by the runtime nor assigned by the file itself."""
    known = KNOWN_GLOBALS | set(extra_known) | set(rep.global_writes)
    return {n: l for n, l in rep.global_reads.items() if n not in known}


def main(argv):
    rc = 0
    for path in argv:
        try:
            rep = check_file(path)
        except LuaSyntaxError as e:
            print(f"SYNTAX  {e}")
            rc = 1
            continue
        unk = unknown_globals(rep)
        for n, l in sorted(unk.items(), key=lambda kv: kv[1]):
            print(f"GLOBAL  {path}:{l}: read of unknown global '{n}'")
            rc = 1
        print(f"ok      {path}: {len(rep.functions)} functions, {rep.locals_declared} locals, "
              f"{len(rep.method_calls)} method calls, global writes: {sorted(rep.global_writes)}")
    return rc


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
