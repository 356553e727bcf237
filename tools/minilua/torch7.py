"""Torch7 stand-in for tools/minilua: `torch.class`, the Tensor type (numpy-backed, 1-based, views share memory) with the
methods of torch7/doc/tensor.md / maths.md that bot7-style code uses, the torch.* constructors and maths functions, and
a minimal `nn` (Sequential / Linear / ReLU / Tanh: forward only) -- enough to EXECUTE lua/bot7_b200/*.lua.

Semantics follow the Torch7 documentation: methods such as add / mul / div / cmul / pow / sqrt / fill / copy work in
place and return the tensor; min / max / sum / mean without a dimension return a number, with a dimension they keep that
dimension with size 1 (min / max also return the LongTensor of 1-based indices); narrow / select / sub / view / expand /
t return views; t[i] selects along the first dimension (a number for 1-D tensors); :data() is a cdata pointer to the
first element (see ffi.py).
"""
import numpy as np

from .interp import LuaError, LuaTable, lua_type, tostring

TYPES = {"torch.DoubleTensor": np.float64, "torch.FloatTensor": np.float32, "torch.LongTensor": np.int64, "torch.IntTensor": np.int32,
         "torch.ByteTensor": np.uint8}
CTYPE = {"torch.DoubleTensor": "double", "torch.FloatTensor": "float", "torch.LongTensor": "int64_t", "torch.IntTensor": "int",
         "torch.ByteTensor": "uint8_t"}


def _i(v, what="index"):
    if isinstance(v, bool) or not isinstance(v, (int, float)) or v != int(v):
        raise LuaError("%s: integer expected, got %s" % (what, tostring(v)))
    return int(v)


def _sizes(args):
    if len(args) == 1 and isinstance(args[0], LongStorage):
        return list(args[0].v)
    if len(args) == 1 and isinstance(args[0], Tensor):         # a LongTensor of sizes
        return [int(x) for x in args[0].a.ravel()]
    if len(args) == 1 and isinstance(args[0], LuaTable):
        return [_i(args[0].get(k), "size") for k in range(1, args[0].length() + 1)]
    return [_i(x, "size") for x in args]


def _table_to_nested(t):
    n = t.length()
    out = []
    for k in range(1, n + 1):
        v = t.get(k)
        out.append(_table_to_nested(v) if isinstance(v, LuaTable) else v)
    return out


class LongStorage:
    """What t:size() / #t return: 1-based, `#` is the number of entries."""
    lua_type = "userdata"

    def __init__(self, values):
        self.v = [int(x) for x in values]

    def lua_index(self, key):
        if isinstance(key, str):
            if key == "totable":
                return lambda self_: LuaTable({k + 1: x for k, x in enumerate(self_.v)})
            if key == "size":
                return lambda self_: len(self_.v)
            return None
        k = _i(key)
        if not 1 <= k <= len(self.v):
            raise LuaError("index out of bounds")
        return self.v[k - 1]

    def lua_len(self):
        return len(self.v)

    def lua_tostring(self):
        return "\n".join(" %d" % x for x in self.v) + "\n[torch.LongStorage of size %d]" % len(self.v)


class TensorStorage:
    """t:storage(): only copy(LongStorage | table) is needed (utils/tensor.lua:shape)."""
    lua_type = "userdata"

    def __init__(self, t):
        self.t = t

    def lua_index(self, key):
        if key == "copy":
            def cp(self_, src):
                vals = src.v if isinstance(src, LongStorage) else (src.a.ravel() if isinstance(src, Tensor) else _table_to_nested(src))
                self_.t.a.ravel()[...] = vals
                return self_
            return cp
        if key == "size":
            return lambda self_: int(self_.t.a.size)
        if isinstance(key, (int, float)):
            return self.t._elem(self.t.a.ravel()[_i(key) - 1])
        return None


class Tensor:
    lua_type = "userdata"

    def __init__(self, a, ttype="torch.DoubleTensor"):
        self.a = a
        self.ttype = ttype

    # ---- protocol
    def lua_tostring(self):
        return "%s\n[%s of size %s]" % (np.array2string(self.a, precision=4), self.ttype, "x".join(str(s) for s in self.a.shape))

    def lua_len(self):
        return LongStorage(self.a.shape)

    def lua_eq(self, other):
        return other is self

    def _wrap(self, a):
        return Tensor(a, self.ttype)

    def _elem(self, v):
        return float(v) if self.a.dtype.kind == "f" else int(v)

    def _ranges(self, key):
        """t[{ {a,b}, {}, k }]"""
        idx = []
        for d in range(1, key.length() + 1):
            r = key.get(d)
            if isinstance(r, LuaTable):
                lo, hi = r.get(1), r.get(2)
                n = self.a.shape[d - 1]
                if lo is None:
                    idx.append(slice(None))
                else:
                    lo = _i(lo)
                    hi = lo if hi is None else _i(hi)
                    lo = n + lo + 1 if lo < 0 else lo
                    hi = n + hi + 1 if hi < 0 else hi
                    if not (1 <= lo <= hi <= n):
                        raise LuaError("index out of bound in dimension %d: {%d,%d} of %d" % (d, lo, hi, n))
                    idx.append(slice(lo - 1, hi))
            elif isinstance(r, Tensor):
                idx.append(r.a.astype(np.int64).ravel() - 1)         # t[{idx}]: a copy of the selected entries (numpy fancy index)
            else:
                k = _i(r)
                n = self.a.shape[d - 1]
                if not 1 <= k <= n:
                    raise LuaError("index %d out of range in dimension %d (size %d)" % (k, d, n))
                idx.append(k - 1)
        return tuple(idx)

    def lua_index(self, key):
        if isinstance(key, str):
            f = getattr(Tensor, "m_" + key, None)
            if f is None:
                return None
            return f
        if isinstance(key, LuaTable):
            r = self.a[self._ranges(key)]
            return self._elem(r) if np.ndim(r) == 0 else self._wrap(r)      # t[{i, j}] with all scalars is a number
        if isinstance(key, Tensor):
            if key.ttype == "torch.ByteTensor":
                return self._wrap(self.a[key.a.astype(bool)])
            raise LuaError("tensor-valued index is not supported by tools/minilua")
        k = _i(key)
        if self.a.ndim == 0:
            raise LuaError("indexing a 0-dimensional tensor")
        if not 1 <= k <= self.a.shape[0]:
            raise LuaError("index %d out of range (size %d)" % (k, self.a.shape[0]))
        if self.a.ndim == 1:
            return self._elem(self.a[k - 1])
        return self._wrap(self.a[k - 1])

    def lua_newindex(self, key, val):
        if isinstance(key, LuaTable):
            tgt = self.a[self._ranges(key)]
            sel = self._ranges(key)
            self.a[sel] = val.a.reshape(np.shape(tgt)) if isinstance(val, Tensor) else val
            return
        k = _i(key)
        if not 1 <= k <= self.a.shape[0]:
            raise LuaError("index %d out of range (size %d)" % (k, self.a.shape[0]))
        if isinstance(val, Tensor):
            self.a[k - 1] = val.a.reshape(self.a[k - 1].shape)
        elif isinstance(val, (int, float)) and not isinstance(val, bool):
            self.a[k - 1] = val
        else:
            raise LuaError("cannot assign a %s to a tensor element" % lua_type(val))

    def lua_arith(self, name, a, b):
        A = a.a if isinstance(a, Tensor) else a
        B = b.a if isinstance(b, Tensor) else b
        tt = a.ttype if isinstance(a, Tensor) else b.ttype
        if name == "__add":
            return Tensor(A + B, tt)
        if name == "__sub":
            return Tensor(A - B, tt)
        if name == "__unm":
            return Tensor(-A, tt)
        if name == "__div":
            if isinstance(b, Tensor):
                raise LuaError("tensor / tensor is not defined in Torch7 (use cdiv)")
            return Tensor(A / B, tt)
        if name == "__mul":
            if isinstance(a, Tensor) and isinstance(b, Tensor):
                r = A @ B
                return float(r) if np.ndim(r) == 0 else Tensor(np.ascontiguousarray(r), tt)
            return Tensor(A * B, tt)
        return NotImplemented

    # ---- shape
    def m_size(self, d=None):
        if d is None:
            return LongStorage(self.a.shape)
        d = _i(d)
        if not 1 <= d <= self.a.ndim:
            raise LuaError("dimension %d out of range of %dD tensor" % (d, self.a.ndim))
        return self.a.shape[d - 1]

    def m_dim(self):
        return self.a.ndim if self.a.size or self.a.ndim > 1 else 0

    m_nDimension = m_dim

    def m_nElement(self):
        return int(self.a.size)

    m_numel = m_nElement

    def m_isContiguous(self):
        return bool(self.a.flags["C_CONTIGUOUS"])

    def m_contiguous(self):
        return self if self.a.flags["C_CONTIGUOUS"] else self._wrap(np.ascontiguousarray(self.a))

    def m_clone(self):
        return self._wrap(np.array(self.a, copy=True, order="C"))

    def m_type(self, t=None):
        if t is None:
            return self.ttype
        if t == self.ttype:
            return self
        if t not in TYPES:
            raise LuaError("unknown tensor type " + tostring(t))
        return Tensor(self.a.astype(TYPES[t]), t)

    def m_typeAs(self, o):
        return self.m_type(o.ttype)

    def m_double(self):
        return self.m_type("torch.DoubleTensor")

    def m_float(self):
        return self.m_type("torch.FloatTensor")

    def m_long(self):
        return self.m_type("torch.LongTensor")

    def m_int(self):
        return self.m_type("torch.IntTensor")

    def m_byte(self):
        return self.m_type("torch.ByteTensor")

    def m_view(self, *sz):
        if not self.a.flags["C_CONTIGUOUS"]:
            raise LuaError("view: expecting a contiguous tensor")
        shape = _sizes(sz)
        if shape.count(-1) > 1:
            raise LuaError("view: only one dimension can be inferred")
        try:
            return self._wrap(self.a.reshape(shape))
        except ValueError:
            raise LuaError("view: size %s is invalid for input of %d elements" % (shape, self.a.size))

    def m_viewAs(self, o):
        return self.m_view(*o.a.shape)

    def m_reshape(self, *sz):
        return self._wrap(np.array(self.a, copy=True).reshape(_sizes(sz)))

    def m_resize(self, *sz):
        shape = _sizes(sz)
        n = int(np.prod(shape)) if shape else 0
        if n == self.a.size and self.a.flags["C_CONTIGUOUS"]:
            self.a = self.a.reshape(shape)
        else:
            new = np.zeros(shape, dtype=self.a.dtype)
            k = min(n, self.a.size)
            new.ravel()[:k] = self.a.ravel()[:k]
            self.a = new
        return self

    def m_resizeAs(self, o):
        return self.m_resize(*o.a.shape)

    def m_narrow(self, dim, index, size):
        dim, index, size = _i(dim), _i(index), _i(size)
        if not 1 <= dim <= self.a.ndim:
            raise LuaError("narrow: dimension %d out of range" % dim)
        if index < 1 or size < 1 or index + size - 1 > self.a.shape[dim - 1]:
            raise LuaError("narrow: out of range (index %d, size %d, dimension size %d)" % (index, size, self.a.shape[dim - 1]))
        sl = [slice(None)] * self.a.ndim
        sl[dim - 1] = slice(index - 1, index - 1 + size)
        return self._wrap(self.a[tuple(sl)])

    def m_sub(self, *r):
        if len(r) % 2 or not r:
            raise LuaError("sub: expects pairs (start, end) per dimension")
        sl = []
        for d in range(len(r) // 2):
            n = self.a.shape[d]
            lo, hi = _i(r[2 * d]), _i(r[2 * d + 1])
            lo = n + lo + 1 if lo < 0 else lo
            hi = n + hi + 1 if hi < 0 else hi
            if not (1 <= lo <= hi <= n):
                raise LuaError("sub: out of range (%d, %d) of %d" % (lo, hi, n))
            sl.append(slice(lo - 1, hi))
        return self._wrap(self.a[tuple(sl)])

    def m_select(self, dim, index):
        dim, index = _i(dim), _i(index)
        if not 1 <= index <= self.a.shape[dim - 1]:
            raise LuaError("select: index out of range")
        sl = [slice(None)] * self.a.ndim
        sl[dim - 1] = index - 1
        r = self.a[tuple(sl)]
        return self._wrap(r)

    def m_t(self):
        if self.a.ndim != 2:
            raise LuaError("t: expecting a 2D tensor")
        return self._wrap(self.a.T)

    def m_transpose(self, d1, d2):
        return self._wrap(np.swapaxes(self.a, _i(d1) - 1, _i(d2) - 1))

    def m_expand(self, *sz):
        shape = _sizes(sz)
        # torch7 (THTensor_expand): at least as many sizes as dimensions; missing leading dimensions are prepended as singletons
        if len(shape) < self.a.ndim:
            raise LuaError("expand: the number of sizes provided must be greater or equal to the number of dimensions in the tensor")
        src = self.a.reshape((1,) * (len(shape) - self.a.ndim) + self.a.shape)
        for s, t in zip(src.shape, shape):
            if s != 1 and s != t:
                raise LuaError("expand: only singleton dimensions can be expanded")
        return self._wrap(np.broadcast_to(src, shape))

    def m_expandAs(self, o):
        return self.m_expand(*o.a.shape)

    def m_repeatTensor(self, *sz):
        return self._wrap(np.tile(self.a, _sizes(sz)))

    def m_squeeze(self, d=None):
        return self._wrap(np.squeeze(self.a) if d is None else (np.squeeze(self.a, _i(d) - 1) if self.a.shape[_i(d) - 1] == 1 else self.a))

    def m_index(self, dim, idx):
        return self._wrap(np.take(self.a, idx.a.astype(np.int64) - 1, axis=_i(dim) - 1))

    def m_data(self):
        from .ffi import CPtr
        if not self.a.flags["C_CONTIGUOUS"]:
            raise LuaError("data(): the tensor is not contiguous (torch would hand out the storage pointer; the C side would read the wrong elements)")
        if not self.a.flags["WRITEABLE"]:
            raise LuaError("data(): expanded tensor")
        return CPtr(CTYPE[self.ttype] + "*", self.a.ctypes.data if self.a.size else 0, keep=self.a)

    # ---- in-place maths
    def _writable(self):
        if not self.a.flags["WRITEABLE"]:
            raise LuaError("in-place operation on an expanded tensor")

    def _other(self, x):
        """The operand of an element-wise in-place operation: TH pairs elements in storage order when the sizes differ but the
        number of elements agrees (numpy would broadcast (M, 1) against (M,) into (M, M))."""
        if not isinstance(x, Tensor):
            return x
        if x.a.shape == self.a.shape:
            return x.a
        if x.a.size == self.a.size:
            return x.a.reshape(self.a.shape)
        raise LuaError("inconsistent tensor size: %s vs %s" % (list(self.a.shape), list(x.a.shape)))

    def _take(self, src):
        """x:op(src, ...): the result goes to x, resized like src."""
        if src is not self:
            if self.a.shape != src.a.shape:
                self.a = np.empty(src.a.shape, dtype=self.a.dtype)
            self.a[...] = src.a
        return self

    def m_fill(self, v):
        self._writable()
        self.a[...] = v
        return self

    def m_zero(self):
        return self.m_fill(0)

    def m_copy(self, o):
        self._writable()
        if not isinstance(o, Tensor):
            raise LuaError("copy: tensor expected, got %s" % lua_type(o))
        if o.a.size != self.a.size:
            raise LuaError("copy: sizes do not match (%s vs %s)" % (self.a.shape, o.a.shape))
        self.a[...] = self._other(o)
        return self

    def m_add(self, *r):
        self._writable()
        if len(r) == 1:                                 # x:add(value) / x:add(tensor)
            self.a += self._other(r[0])
        elif len(r) == 2 and not isinstance(r[0], Tensor):   # x:add(value, tensor)
            self.a += r[0] * self._other(r[1])
        elif len(r) == 2:                                # x:add(tensor1, tensor2 | value)
            self._take(r[0])
            self.a += self._other(r[1])
        elif len(r) == 3:                                # x:add(tensor1, value, tensor2)
            self._take(r[0])
            self.a += r[1] * self._other(r[2])
        else:
            raise LuaError("add: unsupported arguments")
        return self

    def m_csub(self, *r):
        self._writable()
        if len(r) == 1:
            self.a -= self._other(r[0])
        else:
            self.a -= r[0] * self._other(r[1])
        return self

    def m_mul(self, *r):
        self._writable()
        if len(r) == 2:                                  # x:mul(src, value)
            self._take(r[0])
            r = r[1:]
        if isinstance(r[0], Tensor):
            raise LuaError("mul: number expected (use cmul for tensors)")
        self.a *= r[0]
        return self

    def m_div(self, *r):
        self._writable()
        if len(r) == 2:
            self._take(r[0])
            r = r[1:]
        if isinstance(r[0], Tensor):
            raise LuaError("div: number expected (use cdiv for tensors)")
        if self.a.dtype.kind == "f":
            self.a /= r[0]
        else:
            self.a //= int(r[0])
        return self

    def m_cmul(self, *r):
        self._writable()
        if len(r) == 2:
            self._take(r[0])
            r = r[1:]
        self.a *= self._other(r[0])
        return self

    def m_cdiv(self, *r):
        self._writable()
        if len(r) == 2:
            self._take(r[0])
            r = r[1:]
        with np.errstate(all="ignore"):
            self.a /= self._other(r[0])
        return self

    def m_cpow(self, *r):
        self._writable()
        if len(r) == 2:
            self._take(r[0])
            r = r[1:]
        with np.errstate(all="ignore"):
            self.a[...] = self.a ** self._other(r[0])
        return self

    def m_mv(self, m, v):
        self._writable()
        self.a[...] = m.a @ v.a
        return self

    def m_storage(self):
        return TensorStorage(self)

    def m_pow(self, *r):
        self._writable()
        if len(r) == 2:                                  # x:pow(src, value)
            self._take(r[0])
            r = r[1:]
        with np.errstate(all="ignore"):
            self.a[...] = self.a ** r[0]
        return self

    def m_clamp(self, lo, hi):
        self._writable()
        # TH_TENSOR_APPLY: (x < lo) ? lo : ((x > hi) ? hi : x) -- comparison based, a NaN passes through
        x = self.a
        self.a[...] = np.where(x < lo, lo, np.where(x > hi, hi, x))
        return self

    def _unary(f):
        def m(self, src=None):
            self._writable()
            if src is not None:
                self._take(src)
            with np.errstate(all="ignore"):
                self.a[...] = f(self.a)
            return self
        return m

    m_sqrt, m_exp, m_log, m_abs, m_neg, m_floor, m_ceil = (_unary(np.sqrt), _unary(np.exp), _unary(np.log), _unary(np.abs), _unary(np.negative),
                                                          _unary(np.floor), _unary(np.ceil))
    m_cos, m_sin, m_tan, m_tanh, m_round = _unary(np.cos), _unary(np.sin), _unary(np.tan), _unary(np.tanh), _unary(np.round)
    del _unary

    def m_apply(self, f):
        flat = self.a.ravel() if self.a.flags["C_CONTIGUOUS"] else None
        if flat is None:
            raise LuaError("apply on a non-contiguous tensor is not supported by tools/minilua")
        for k in range(flat.size):
            r = f(self._elem(flat[k]))
            r = r[0] if isinstance(r, list) and r else r
            if r is not None and not isinstance(r, list):
                flat[k] = r
        return self

    # ---- reductions
    def _reduce(self, f, d, with_index=None):
        if self.a.size == 0:
            raise LuaError("reduction of an empty tensor")
        if d is None:
            return self._elem(f(self.a))
        ax = _i(d) - 1
        if not 0 <= ax < self.a.ndim:
            raise LuaError("dimension %d out of range" % (ax + 1))
        vals = self._wrap(np.ascontiguousarray(f(self.a, axis=ax, keepdims=True)))
        if with_index is None:
            return vals
        idx = Tensor(np.ascontiguousarray(np.expand_dims(with_index(self.a, axis=ax), ax).astype(np.int64) + 1), "torch.LongTensor")
        return [vals, idx]

    def m_min(self, d=None):
        return self._reduce(np.min, d, np.argmin)

    def m_max(self, d=None):
        return self._reduce(np.max, d, np.argmax)

    def m_sum(self, d=None):
        return self._reduce(np.sum, d)

    def m_mean(self, d=None):
        if d is None:
            return float(np.mean(self.a))
        return self._reduce(np.mean, d)

    def m_prod(self, d=None):
        return self._reduce(np.prod, d)

    def m_std(self, *r):
        return float(np.std(self.a, ddof=1))

    def m_var(self, *r):
        return float(np.var(self.a, ddof=1))

    def m_norm(self, p=2):
        if p == 2:
            return float(np.sqrt(np.sum(self.a * self.a)))           # TH: sqrt of the sum of squares
        return float(np.sum(np.abs(self.a) ** p) ** (1.0 / p))

    def m_dot(self, o):
        return float(np.dot(self.a.ravel(), o.a.ravel()))

    def _cmp(f):
        def m(self, *r):
            if len(r) == 2:                              # x:ge(src, value): 0 / 1 into x, keeping x's type
                src, o = r
                res = f(src.a, o.a if isinstance(o, Tensor) else o)
                if self.a.shape != src.a.shape:
                    self.a = np.empty(src.a.shape, dtype=self.a.dtype)
                self.a[...] = res
                return self
            o = r[0]
            return Tensor(f(self.a, self._other(o) if isinstance(o, Tensor) else o).astype(np.uint8), "torch.ByteTensor")
        return m

    m_eq, m_ne, m_lt, m_le, m_gt, m_ge = _cmp(np.equal), _cmp(np.not_equal), _cmp(np.less), _cmp(np.less_equal), _cmp(np.greater), _cmp(np.greater_equal)
    del _cmp

    def m_maskedFill(self, mask, val):
        self._writable()
        self.a[mask.a.astype(bool).reshape(self.a.shape)] = val
        return self

    def m_maskedCopy(self, mask, src):
        self._writable()
        m = mask.a.astype(bool).reshape(self.a.shape)
        n = int(m.sum())
        if src.a.size < n:
            raise LuaError("maskedCopy: the source has fewer elements than the mask selects")
        self.a[m] = src.a.ravel()[:n]
        return self

    def m_maskedSelect(self, mask):
        return self._wrap(np.ascontiguousarray(self.a[mask.a.astype(bool).reshape(self.a.shape)]))

    def m_cat(self, other, dim=None):
        d = _i(dim) if dim is not None else self.a.ndim
        return self._wrap(np.concatenate([self.a, other.a], axis=d - 1))

    def m_any(self):
        return bool(np.any(self.a != 0))

    def m_all(self):
        return bool(np.all(self.a != 0))

    def m_nonzero(self):
        idx = np.argwhere(self.a != 0).astype(np.int64) + 1
        if idx.shape[0] == 0:
            return Tensor(np.zeros((0,), dtype=np.int64), "torch.LongTensor")       # dim() == 0, as torch7's empty result
        return Tensor(np.ascontiguousarray(idx), "torch.LongTensor")

    def m_indexCopy(self, dim, idx, src):
        self._writable()
        ax = _i(dim) - 1
        k = idx.a.astype(np.int64).ravel() - 1
        sl = [slice(None)] * self.a.ndim
        sl[ax] = k
        self.a[tuple(sl)] = src.a
        return self

    def m_indexFill(self, dim, idx, val):
        self._writable()
        sl = [slice(None)] * self.a.ndim
        sl[_i(dim) - 1] = idx.a.astype(np.int64).ravel() - 1
        self.a[tuple(sl)] = val
        return self

    def m_totable(self):
        def conv(a):
            t = LuaTable()
            for k, v in enumerate(a):
                t.set(k + 1, conv(v) if np.ndim(v) else self._elem(v))
            return t
        return conv(self.a)


def make_tensor(ttype, args, rng=None):
    dt = TYPES[ttype]
    if not args:
        return Tensor(np.zeros((0,), dtype=dt), ttype)
    if len(args) == 1 and isinstance(args[0], LuaTable):
        data = np.array(_table_to_nested(args[0]), dtype=dt)
        return Tensor(np.ascontiguousarray(data), ttype)
    if len(args) == 1 and isinstance(args[0], Tensor):
        return Tensor(args[0].a if args[0].ttype == ttype else args[0].a.astype(dt), ttype)     # shares the storage
    # torch.Tensor(sizes...) is uninitialised memory: poison it, so that a glue that reads before it writes is caught
    a = np.empty(_sizes(args), dtype=dt)
    a[...] = np.nan if dt in (np.float64, np.float32) else (-(2 ** 30) if dt in (np.int64, np.int32) else 171)
    return Tensor(a, ttype)


def install(I, seed=0):
    """Puts `torch` (and `nn`) into the interpreter's globals."""
    T = LuaTable()
    state = {"rng": np.random.default_rng(seed), "default": "torch.DoubleTensor"}
    classes = {}

    def typename(v):
        if isinstance(v, Tensor):
            return v.ttype
        if isinstance(v, LuaTable) and v.meta is not None:
            n = v.meta.get("__typename")
            if n is not None:
                return n
        return None

    def t_type(v):
        return typename(v) or lua_type(v)

    def t_class(name, parent_name=None):
        name = tostring(name)
        parent = None
        if parent_name is not None:
            parent = classes.get(tostring(parent_name))
            if parent is None:
                raise LuaError("torch.class: parent class '%s' is not defined" % tostring(parent_name))
        cls = LuaTable()
        objmeta = LuaTable()
        objmeta.set("__index", cls)
        objmeta.set("__typename", name)

        def obj_call(self, *a):
            f = I.index(self, "__call__")
            if f is None:
                raise LuaError("attempt to call an object of class %s, which defines no __call__" % name)
            return I.call(f, [self] + list(a))
        objmeta.set("__call", obj_call)

        def obj_tostring(self):
            f = I.index(self, "__tostring__")
            if f is not None:
                r = I.call(f, [self])
                return r[0] if r else "nil"
            return name
        objmeta.set("__tostring", obj_tostring)

        def ctor(_cls, *a):
            obj = LuaTable(meta=objmeta)
            init = I.index(cls, "__init")
            if init is not None:
                I.call(init, [obj] + list(a))
            return obj
        clsmeta = LuaTable()
        clsmeta.set("__call", ctor)
        if parent is not None:
            clsmeta.set("__index", parent)
        cls.meta = clsmeta
        classes[name] = cls
        # luaT_lua_newmetatable / luaT_getinnerparent (torch7/lib/luaT/luaT.c): a dotted name stores the class in the EXISTING
        # nested global tables of its package part and fails when one of them is missing -- the well-known
        # "while creating metatable a.b.C: bad argument #1 (a is an invalid module name)"
        parts = name.split(".")
        tbl = I.G
        for p in parts[:-1]:
            nxt = tbl.get(p)
            if not isinstance(nxt, LuaTable):
                raise LuaError("while creating metatable %s: bad argument #1 (%s is an invalid module name)" % (name, p))
            tbl = nxt
        tbl.set(parts[-1], cls)
        return [cls, parent]

    def ctor_of(tt):
        return lambda *a: make_tensor(tt, a)

    def new_filled(v):
        def f(*sz):
            return Tensor(np.full(_sizes(sz), v, dtype=TYPES[state["default"]]), state["default"])
        return f

    def t_rand(*sz):
        return Tensor(state["rng"].random(_sizes(sz)), "torch.DoubleTensor")

    def t_randn(*sz):
        return Tensor(state["rng"].standard_normal(_sizes(sz)), "torch.DoubleTensor")

    def t_randperm(n):
        return Tensor((state["rng"].permutation(_i(n)) + 1).astype(np.float64), "torch.DoubleTensor")

    def t_seed(s=0):
        state["rng"] = np.random.default_rng(_i(s))

    class RNGState:
        lua_type = "userdata"

        def __init__(self, st):
            self.st = st

    def t_get_state():
        import copy
        return RNGState(copy.deepcopy(state["rng"].bit_generator.state))

    def t_set_state(st):
        import copy
        if not isinstance(st, RNGState):
            raise LuaError("setRNGState: a state returned by torch.getRNGState() expected")
        state["rng"].bit_generator.state = copy.deepcopy(st.st)

    def t_cat(*a):
        if isinstance(a[0], LuaTable):
            ts = [a[0].get(k) for k in range(1, a[0].length() + 1)]
            dim = a[1] if len(a) > 1 else None
        else:
            ts = [x for x in a if isinstance(x, Tensor)]
            dim = a[len(ts)] if len(a) > len(ts) else None
        if not ts:
            raise LuaError("cat: empty list")
        dim = _i(dim) if dim is not None else ts[0].a.ndim
        return Tensor(np.concatenate([t.a for t in ts], axis=dim - 1), ts[0].ttype)

    def fresh(f):
        """torch.f(tensor, ...) -> new tensor (the method works in place on a clone); torch.f(result, tensor, ...) fills result."""
        def g(*a):
            if len(a) >= 2 and isinstance(a[0], Tensor) and isinstance(a[1], Tensor) and f in ("sqrt", "exp", "log", "abs", "neg", "pow", "mul", "div", "cos", "sin", "tanh", "floor", "ceil"):
                res, src, rest = a[0], a[1], a[2:]
                res.m_resizeAs(src).m_copy(src)
                getattr(Tensor, "m_" + f)(res, *rest)
                return res
            c = a[0].m_clone()
            return getattr(Tensor, "m_" + f)(c, *a[1:])
        return g

    def t_add(*a):
        if isinstance(a[0], Tensor):
            return a[0].m_clone().m_add(*a[1:])
        return a[1].m_clone().m_add(a[0])

    def t_range(lo, hi, step=1, ttype=None):
        if isinstance(step, str):                       # torch.range(a, b, 'torch.LongTensor')
            step, ttype = 1, step
        tt = ttype or "torch.DoubleTensor"
        n = int(np.floor((hi - lo) / step + 1e-12)) + 1
        return Tensor((lo + step * np.arange(max(n, 0))).astype(TYPES[tt]), tt)

    def t_mm(a, b):
        return Tensor(np.ascontiguousarray(a.a @ b.a), a.ttype)

    for tt, short in [("torch.DoubleTensor", "DoubleTensor"), ("torch.FloatTensor", "FloatTensor"), ("torch.LongTensor", "LongTensor"),
                      ("torch.IntTensor", "IntTensor"), ("torch.ByteTensor", "ByteTensor")]:
        T.set(short, ctor_of(tt))
    T.set("Tensor", lambda *a: make_tensor(state["default"], a))
    for name, f in [("class", t_class), ("type", t_type), ("typename", lambda v: [typename(v)]), ("isTensor", lambda v=None: isinstance(v, Tensor)),
                    ("zeros", new_filled(0)), ("ones", new_filled(1)), ("rand", t_rand), ("randn", t_randn), ("randperm", t_randperm),
                    ("manualSeed", t_seed), ("getRNGState", t_get_state), ("setRNGState", t_set_state), ("cat", t_cat), ("add", t_add), ("range", t_range), ("mm", t_mm),
                    ("cmul", lambda a, b: a.m_clone().m_cmul(b)), ("cdiv", lambda a, b: a.m_clone().m_cdiv(b)),
                    ("mul", fresh("mul")), ("div", fresh("div")), ("sqrt", fresh("sqrt")), ("exp", fresh("exp")), ("log", fresh("log")),
                    ("abs", fresh("abs")), ("pow", fresh("pow")), ("neg", fresh("neg")), ("floor", fresh("floor")), ("ceil", fresh("ceil")),
                    ("cos", fresh("cos")), ("sin", fresh("sin")), ("tanh", fresh("tanh")),
                    ("mv", lambda m, v: Tensor(np.ascontiguousarray(m.a @ v.a), m.ttype)),
                    ("min", lambda t, d=None: t.m_min(d)), ("max", lambda t, d=None: t.m_max(d)), ("sum", lambda t, d=None: t.m_sum(d)),
                    ("mean", lambda t, d=None: t.m_mean(d)), ("dot", lambda a, b: a.m_dot(b)), ("norm", lambda t, p=2: t.m_norm(p)),
                    ("eye", lambda n: Tensor(np.eye(_i(n)), "torch.DoubleTensor")),
                    ("setdefaulttensortype", lambda t: state.__setitem__("default", tostring(t))),
                    ("getdefaulttensortype", lambda: state["default"])]:
        T.set(name, f)
    I.G.set("torch", T)
    I.G.get("package").get("loaded").set("torch", T)

    # ---- nn: forward-only Linear / ReLU / Tanh / Sequential, as objects of torch.class so that torch.type() names them
    I.run(_NN_LUA, "=nn (tools/minilua)")
    return T


_NN_LUA = r"""
nn = nn or {}
do
  local Module = torch.class('nn.Module')
  function Module:__init() self.output = torch.Tensor(); self.train = true end
  function Module:forward(input) return self:updateOutput(input) end
  function Module:evaluate() self.train = false; return self end
  function Module:training() self.train = true; return self end

  local Linear, parent = torch.class('nn.Linear', 'nn.Module')
  function Linear:__init(nIn, nOut)
    parent.__init(self)
    self.weight = torch.randn(nOut, nIn):mul(1 / math.sqrt(nIn))
    self.bias   = torch.randn(nOut):mul(0.1)
  end
  function Linear:updateOutput(input)
    local x = input
    if x:dim() == 1 then x = x:view(1, -1) end
    self.output = torch.mm(x, self.weight:t()):add(self.bias:view(1, -1):expand(x:size(1), self.weight:size(1)))
    return self.output
  end

  local ReLU, parentR = torch.class('nn.ReLU', 'nn.Module')
  function ReLU:__init() parentR.__init(self) end
  function ReLU:updateOutput(input)
    self.output = input:clone():apply(function(v) if v < 0 then return 0 end end)
    return self.output
  end

  local Tanh, parentT = torch.class('nn.Tanh', 'nn.Module')
  function Tanh:__init() parentT.__init(self) end
  function Tanh:updateOutput(input)
    self.output = input:clone():apply(function(v) return math.tanh(v) end)
    return self.output
  end

  local Seq, parentS = torch.class('nn.Sequential', 'nn.Module')
  function Seq:__init() parentS.__init(self); self.modules = {} end
  function Seq:add(m) self.modules[#self.modules + 1] = m; return self end
  function Seq:get(i) return self.modules[i] end
  function Seq:size() return #self.modules end
  function Seq:evaluate() for _, m in ipairs(self.modules) do m:evaluate() end; self.train = false; return self end
  function Seq:updateOutput(input)
    local x = input
    for _, m in ipairs(self.modules) do x = m:forward(x) end
    self.output = x
    return x
  end
end
"""
