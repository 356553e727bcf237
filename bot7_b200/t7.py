"""Torch7 binary serialisation -- the format `torch.save` / `torch.load` write by default.

bot7 persists an experiment as `torch.save('demo_<class>.t7', {best=..., x=observed, y=responses})`
(reference bots/abstract.lua:234-240) and reads data sets the same way (examples/autoML.lua:57,
examples/data/*.t7).  This module reads and writes that format so that a run started on the
reference can be resumed here through the `cache` protocol (bots/abstract.lua:19-44) and the other
way round.  Torch7 itself is not vendored in the reference; the
format below is torch7's `File:writeObject` / `File:readObject`, pinned by the reference's own
examples/data fixtures (tests/test_t7.py).

Stream (little endian; int = 4 bytes, long = 8 bytes, double = 8 bytes):
    object  := int type, payload
    NIL 0   := -
    NUMBER 1:= double
    STRING 2:= int n, n bytes
    BOOLEAN 5 := int 0/1
    TABLE 3 := int index, [int npairs, npairs x (object key, object value)]      (body only on first sight)
    TORCH 4 := int index, [string version "V 1", string class, class body]       (body only on first sight)
    tensor body  := int ndim, long size[ndim], long stride[ndim], long offset (1-based), object storage
    storage body := long n, n raw elements
Indices number tables, tensors and storages in the order they are first written; a repeated index
refers to the object already seen (shared references survive a round trip of the reader).

Lua tables map to `dict` (keys in file order; integral number keys become `int`), tensors to numpy
arrays (strided views of their storage), storages to 1-d numpy arrays.  The writer maps `dict` to a
table, `list` / `tuple` to a 1-based array table, numpy arrays to tensors with a private storage.
"""
from __future__ import annotations

import io
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5

_ELEM = {"Double": np.float64, "Float": np.float32, "Long": np.int64, "Int": np.int32, "Short": np.int16,
         "Char": np.int8, "Byte": np.uint8}
_NAME = {np.dtype(v): k for k, v in _ELEM.items()}


class T7Error(ValueError):
    pass


# ---- reader ------------------------------------------------------------------------------------

class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        self.p = 0
        self.memo = {}

    def _take(self, n):
        if self.p + n > len(self.d):
            raise T7Error("truncated .t7 stream at byte %d" % self.p)
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def int(self):
        return struct.unpack("<i", self._take(4))[0]

    def long(self):
        return struct.unpack("<q", self._take(8))[0]

    def longs(self, n):
        return list(struct.unpack("<%dq" % n, self._take(8 * n))) if n else []

    def string(self):
        return self._take(self.int()).decode("latin-1")

    def object(self):
        t = self.int()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            return struct.unpack("<d", self._take(8))[0]
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_BOOLEAN:
            return self.int() != 0
        if t == TYPE_TABLE:
            idx = self.int()
            if idx in self.memo:
                return self.memo[idx]
            out = self.memo[idx] = {}
            for _ in range(self.int()):
                k = self.object()
                v = self.object()
                if isinstance(k, float) and k.is_integer():
                    k = int(k)
                out[k] = v
            return out
        if t == TYPE_TORCH:
            idx = self.int()
            if idx in self.memo:
                return self.memo[idx]
            version = self.string()
            cls = self.string() if version.startswith("V ") else version     # pre-versioning files: the string is the class
            obj = self.memo[idx] = self.torch_object(cls)
            return obj
        raise T7Error("unsupported .t7 type tag %d at byte %d (functions and userdata are not data)" % (t, self.p - 4))

    def torch_object(self, cls):
        if not cls.startswith("torch."):
            raise T7Error("unsupported torch class %r" % cls)
        name = cls[len("torch."):]
        if name.endswith("Storage") and name[:-7] in _ELEM:
            dt = np.dtype(_ELEM[name[:-7]])
            n = self.long()
            return np.frombuffer(self._take(n * dt.itemsize), dtype=dt).copy()
        if name.endswith("Tensor") and name[:-6] in _ELEM:
            dt = np.dtype(_ELEM[name[:-6]])
            nd = self.int()
            size, stride = self.longs(nd), self.longs(nd)
            off = self.long() - 1
            storage = self.object()
            if nd == 0 or storage is None:
                return np.empty((0,), dtype=dt)
            if storage.dtype != dt:
                raise T7Error("tensor %s over a %s storage" % (cls, storage.dtype))
            need = off + sum((s - 1) * st for s, st in zip(size, stride)) + 1 if all(s > 0 for s in size) else 0
            if off < 0 or need > storage.size:
                raise T7Error("tensor view exceeds its storage")
            return np.lib.stride_tricks.as_strided(storage[off:], shape=size, strides=[st * dt.itemsize for st in stride], writeable=False)
        raise T7Error("unsupported torch class %r" % cls)


def loads(data: bytes):
    r = _Reader(data)
    obj = r.object()
    if r.p != len(data):
        raise T7Error("%d trailing bytes after the root object" % (len(data) - r.p))
    return obj


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())


# ---- writer ------------------------------------------------------------------------------------

class _Writer:
    def __init__(self):
        self.b = io.BytesIO()
        self.count = 0
        self.memo = {}        # id(obj) -> index (tables only: arrays always get a fresh storage)
        self.keep = []

    def int(self, v):
        self.b.write(struct.pack("<i", v))

    def long(self, v):
        self.b.write(struct.pack("<q", v))

    def string(self, s):
        raw = s.encode("latin-1")
        self.int(len(raw))
        self.b.write(raw)

    def new_index(self):
        self.count += 1
        self.int(self.count)

    def object(self, o):
        if o is None:
            self.int(TYPE_NIL)
        elif isinstance(o, (bool, np.bool_)):
            self.int(TYPE_BOOLEAN)
            self.int(1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.int(TYPE_NUMBER)
            self.b.write(struct.pack("<d", float(o)))
        elif isinstance(o, str):
            self.int(TYPE_STRING)
            self.string(o)
        elif isinstance(o, np.ndarray):
            self.tensor(o)
        elif isinstance(o, (dict, list, tuple)):
            self.int(TYPE_TABLE)
            if id(o) in self.memo:
                self.int(self.memo[id(o)])
                return
            self.new_index()
            self.memo[id(o)] = self.count
            self.keep.append(o)
            items = list(o.items()) if isinstance(o, dict) else list(enumerate(o, 1))
            self.int(len(items))
            for k, v in items:
                self.object(k)
                self.object(v)
        else:
            raise T7Error("cannot serialise %r as Torch7 data" % type(o))

    def tensor(self, a):
        if a.dtype not in _NAME:
            raise T7Error("no Torch7 tensor type for dtype %s" % a.dtype)
        name = _NAME[a.dtype]
        self.int(TYPE_TORCH)
        self.new_index()
        self.string("V 1")
        self.string("torch.%sTensor" % name)
        if a.size == 0:
            self.int(0)
            self.long(1)
            self.int(TYPE_NIL)
            return
        c = np.ascontiguousarray(a)
        if c.ndim == 0:
            c = c.reshape(1)
        self.int(c.ndim)
        for s in c.shape:
            self.long(s)
        for st in c.strides:
            self.long(st // c.itemsize)
        self.long(1)
        self.int(TYPE_TORCH)
        self.new_index()
        self.string("V 1")
        self.string("torch.%sStorage" % name)
        self.long(c.size)
        self.b.write(c.astype(c.dtype.newbyteorder("<"), copy=False).tobytes())


def dumps(obj) -> bytes:
    w = _Writer()
    w.object(obj)
    return w.b.getvalue()


def save(path, obj):
    data = dumps(obj)
    with open(path, "wb") as f:
        f.write(data)
