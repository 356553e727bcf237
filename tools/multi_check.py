"""One process per GPU (torchrun): the multi-GPU block of the C ABI against the one-GPU result on rank 0.
Run: python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/multi_check.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch.distributed as dist  # noqa: E402

import b7_oracle as o  # noqa: E402
from bot7_b200 import _lib as L  # noqa: E402
from bot7_b200 import grids, models, parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
comm = parallel.Comm.from_env(local, world, rank)
N, d, M = 700, 6, 30011
pts = o.sobol_points(d, N + M)
Xo = pts[:N]
y = o.hartmann6(Xo)
y = (y - y.mean()) / y.std()
for S in (8, 5):
    r = np.random.default_rng(S)
    hyp = np.zeros((S, d + 3))
    hyp[:, :d] = np.log(0.15) + r.random((S, d)) * np.log(8)
    hyp[:, d + 1] = 0.5 * np.log(1e-2)
    gs = comm.sobol_grid(d, 1 + N, M)
    gps, info, logml, jit, gather_ms = comm.fit(Xo, y, hyp)
    b, a, ao, n_, sc = comm.acq_score(gps, gs, L.SCORE_EI, 0.0, 0, -1.0, float(y.min()), want_scores=True)
    row = comm.grid_remove(gs, a)
    b2, a2, ao2, n2, _ = comm.acq_score(gps, gs, L.SCORE_EI, 0.0, 0, -1.0, float(y.min()))
    if rank == 0:
        ctx = comm.ctxs[0]
        f = models.GPFactors(Xo, y, hyp, ctx=ctx)
        g1 = grids.sobol({"size": N + M, "dims": d}, ctx=ctx).generate_device(first=N, count=M)
        one = np.empty(M)
        am, amo, best, nn = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
        L.check(L.lib().b7_acq_score(f.handle, g1.handle, L.SCORE_EI, 0.0, 0, -1.0, float(y.min()), L.dptr(one), C.byref(am), C.byref(amo),
                                     C.byref(best), C.byref(nn)))
        r0, cnt = parallel.shard_range(M, world, 0)
        assert np.array_equal(logml, f.logml), "log marginal likelihoods differ"
        assert np.array_equal(sc, one[r0:r0 + cnt]), "shard scores differ"
        assert (b, a, ao, n_) == (best.value, am.value, amo.value, nn.value), ((b, a, ao, n_), (best.value, am.value))
        row1 = g1.remove(am.value)
        assert np.array_equal(row, row1.reshape(-1)), "removed rows differ"
        L.check(L.lib().b7_acq_score(f.handle, g1.handle, L.SCORE_EI, 0.0, 0, -1.0, float(y.min()), None, C.byref(am), C.byref(amo),
                                     C.byref(best), C.byref(nn)))
        assert (b2, a2, ao2, n2) == (best.value, am.value, amo.value, nn.value)
        f.free()
        g1.free()
        print(f"S={S}: world {world}, gather {gather_ms:.3f} ms, argmax {a} ok")
    comm.free_fit(gps)
    for g in gs:
        g.free()
dist.barrier()
if rank == 0:
    print("multi_check ok")
comm.close()
dist.destroy_process_group()
