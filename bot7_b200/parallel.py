"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards by candidates: rank g owns the contiguous original-row range
shard_range(M, G, g) of the Sobol sequence and generates it locally.  The only exchanges are
  * fit:     draws are factorised S/G per rank, then L^-1 and beta are all-gathered (NCCL over
             NVLink; in place on the library's own device buffers), or every rank factorises all
             draws when S < G;
  * combine: an all-gather of one (best score, global original index, nan count) triple per rank,
             then every rank applies the same rule -- larger score wins, ties go to the smaller
             global index -- which reproduces the reference's first-maximum scan
             (bots/bayesopt.lua:96) for any G.
"""
from __future__ import annotations

import math


def shard_range(M: int, world: int, rank: int):
    """Contiguous shard [row0, row0+count) of M rows; the first M % world ranks get one extra row."""
    base, extra = divmod(int(M), int(world))
    row0 = rank * base + min(rank, extra)
    return row0, base + (1 if rank < extra else 0)


def draw_range(S: int, world: int, rank: int):
    return shard_range(S, world, rank)


def combine_argmax(triples):
    """triples: iterable of (best, global_index_1based or 0, nan_count).  Deterministic reduction:
    NaN/empty shards (index 0) are skipped; max score, ties -> smallest index."""
    best, idx, nans = float("nan"), 0, 0
    for b, i, n in triples:
        nans += int(n)
        if int(i) <= 0 or (isinstance(b, float) and math.isnan(b)):
            continue
        if idx == 0 or b > best or (b == best and int(i) < idx):
            best, idx = float(b), int(i)
    return best, idx, nans


def allgather_argmax(best, idx, nans, group=None):
    """All ranks get the same (best, idx, nans).  CPU (gloo) and CUDA (NCCL) tensors both work."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return combine_argmax([(best, idx, nans)])
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    world = dist.get_world_size(group)
    # index and count travel as exact int64; the score as float64
    f = torch.tensor([float(best)], dtype=torch.float64, device=dev)
    i = torch.tensor([int(idx), int(nans)], dtype=torch.int64, device=dev)
    fs = [torch.empty_like(f) for _ in range(world)]
    is_ = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(fs, f, group=group)
    dist.all_gather(is_, i, group=group)
    return combine_argmax([(float(a[0]), int(b[0]), int(b[1])) for a, b in zip(fs, is_)])


class _DevBuf:
    """__cuda_array_interface__ view of a raw device pointer so torch can wrap library memory."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def allgather_draws(full, per_draw, S, world, rank, group=None):
    """In-place all-gather of a flat per-draw buffer: rank r owns draws draw_range(S, world, r) of `full`
    (S * per_draw elements) and receives everybody else's.  Works on CPU (gloo) and CUDA (NCCL) tensors."""
    import torch.distributed as dist
    chunks = []
    for r in range(world):
        s0, cnt = draw_range(S, world, r)
        chunks.append(full[s0 * per_draw:(s0 + cnt) * per_draw])
    if all(c.numel() == chunks[0].numel() for c in chunks):
        dist.all_gather(chunks, chunks[rank].clone(), group=group)
    else:                                   # S not divisible by the world size: one broadcast per owner
        for r in range(world):
            if chunks[r].numel():
                dist.broadcast(chunks[r], src=r, group=group)
    return full


def allgather_factors(factors, S, world, rank, group=None):
    """After each rank factorised + inverted its draw range: all-gather L^-1 (S x Np x Np, tiled layout) and
    beta (S x Np) in place on the library's own device buffers (NCCL over NVLink)."""
    import torch
    Np = factors.padded_n()
    for what, per_draw in ((0, Np * Np), (1, Np)):
        ptr, nbytes = factors.device_ptr(what)
        full = torch.as_tensor(_DevBuf(ptr, nbytes), device="cuda")
        torch.cuda.synchronize()
        allgather_draws(full, per_draw, S, world, rank, group)
        torch.cuda.synchronize()
