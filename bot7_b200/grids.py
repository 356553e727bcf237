"""bot7.grids -- candidate grids (host-side mirror of reference grids/*.lua over the C ABI).

`sobol(config)` / `random(config)` keep the reference's constructor + call protocol
(grids/abstract.lua:20-23): `Grid(config)()` returns a size x dims fp64 array.  The points are
generated on the GPU by b7_sobol_generate; `generate_device()` additionally keeps them resident
(a DeviceGrid) for bots.bayesopt, which never needs the grid on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class DeviceGrid:
    """Device-resident candidate grid with the reference's compaction semantics
    (utils.tensor.steal / remove, utils/tensor.lua:158-193): rows keep their original position on the
    device; removed rows are tombstones and indices seen by the caller are the compacted ones."""

    def __init__(self, handle, ctx):
        self.handle, self.ctx = handle, ctx

    @classmethod
    def from_host(cls, X, ctx=None):
        ctx = ctx or L.Context.default()
        X = L.as_f64(X)
        if X.ndim != 2:
            raise ValueError("grid must be M x d")
        h = C.c_void_p()
        L.check(L.lib().b7_grid_from_host(ctx.handle, L.dptr(X), X.shape[0], X.shape[1], C.byref(h)), "grid_from_host")
        return cls(h, ctx)

    def size(self) -> int:
        return int(L.lib().b7_grid_size(self.handle))

    def rows(self) -> int:
        return int(L.lib().b7_grid_rows(self.handle))

    def dims(self) -> int:
        return int(L.lib().b7_grid_dims(self.handle))

    def read(self, first=0, count=None) -> np.ndarray:
        count = self.rows() - first if count is None else count
        out = np.empty((count, self.dims()))
        L.check(L.lib().b7_grid_read(self.handle, first, count, L.dptr(out)), "grid_read")
        return out

    def remove(self, compacted_index_1based: int) -> np.ndarray:
        """steal(): returns the removed row (1 x d)."""
        row = np.empty((1, self.dims()))
        L.check(L.lib().b7_grid_remove(self.handle, int(compacted_index_1based), L.dptr(row)), "grid_remove")
        return row

    def original_index(self, compacted_index_1based: int) -> int:
        out = C.c_int64()
        L.check(L.lib().b7_grid_original_index(self.handle, int(compacted_index_1based), C.byref(out)))
        return out.value

    def free(self):
        if self.handle:
            L.lib().b7_grid_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _row(v, dims):
    return None if v is None else L.as_f64(v).reshape(-1)[:dims].copy()


def _rescale_one_sided(g, mins, maxes):
    # grids/sobol.lua:82-86 (one-sided variants; the bots always pass both bounds)
    if mins is not None:
        return g + (mins.reshape(1, -1) + g.min(axis=0, keepdims=True))
    return g * (maxes.reshape(1, -1) / g.max(axis=0, keepdims=True))


class sobol:
    """bot7.grids.sobol (grids/sobol.lua:27-90)."""

    def __init__(self, config=None, ctx=None):
        c = dict(config or {})
        c.setdefault("max_dims", 40)       # grids/sobol.lua:31
        c.setdefault("log_max", 30)        # :32
        c.setdefault("bit_precis", 32)     # :33
        assert c.get("size") is not None   # :35
        assert c.get("dims") is not None and c["dims"] < c["max_dims"]   # :36
        self.config = c
        self.ctx = ctx

    def _args(self, config):
        c = self.config if config is None else config
        dims, size = int(c["dims"]), int(c["size"])
        skip = c.get("skip")
        skip = 1 if skip is None else int(skip)                # :70
        return c, dims, size, skip, _row(c.get("mins"), dims), _row(c.get("maxes"), dims)

    def generate(self, config=None) -> np.ndarray:
        c, dims, size, skip, mins, maxes = self._args(config)
        ctx = self.ctx or L.Context.default()
        out = np.empty((size, dims))
        both = mins is not None and maxes is not None
        L.check(L.lib().b7_sobol_generate(ctx.handle, dims, skip, size, L.dptr(mins) if both else None,
                                          L.dptr(maxes) if both else None, L.dptr(out), None), "sobol_generate")
        if not both and (mins is not None or maxes is not None):
            out = _rescale_one_sided(out, mins, maxes)
        return out

    def generate_device(self, config=None, first=None, count=None) -> DeviceGrid:
        """Same points kept on the GPU; (first, count) selects a sub-range of the sequence so that each
        rank of a multi-GPU run generates only its own shard."""
        c, dims, size, skip, mins, maxes = self._args(config)
        if (mins is None) != (maxes is None):
            raise ValueError("device grids need both mins and maxes (or neither)")
        ctx = self.ctx or L.Context.default()
        first = 0 if first is None else first
        count = size - first if count is None else count
        h = C.c_void_p()
        L.check(L.lib().b7_sobol_generate(ctx.handle, dims, skip + first, count, L.dptr(mins), L.dptr(maxes), None,
                                          C.byref(h)), "sobol_generate")
        return DeviceGrid(h, ctx)

    __call__ = generate


class random:
    """bot7.grids.random (grids/random.lua:18-35): host RNG (torch.rand there, numpy here -- the RNG
    stream is host state and is not reproduced), same affine map."""

    def __init__(self, config=None, rng=None):
        self.config = dict(config or {})
        self.rng = rng or np.random.default_rng()

    def generate(self, config=None) -> np.ndarray:
        c = self.config if config is None else config
        dims, size = int(c["dims"]), int(c["size"])
        g = self.rng.random((size, dims))
        mins, maxes = _row(c.get("mins"), dims), _row(c.get("maxes"), dims)
        if mins is not None and maxes is not None:
            g = g * (maxes + (-mins)).reshape(1, -1)
            g = g + mins.reshape(1, -1)
        elif mins is not None or maxes is not None:
            g = _rescale_one_sided(g, mins, maxes)
        return g

    def generate_device(self, config=None, ctx=None) -> DeviceGrid:
        return DeviceGrid.from_host(self.generate(config), ctx)

    __call__ = generate
