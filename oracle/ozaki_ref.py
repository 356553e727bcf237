"""CPU restatement of the error-free int8 slicing used by the INT8 tensor path (TEST INFRASTRUCTURE ONLY).

Not a reference-repository algorithm: it restates, in exact integer arithmetic, what
bot7_b200/csrc/i8_common.cuh (digit_bytes, pack4), posterior_i8.cu, potrf_i8.cu and trtri_i8.cu compute, so
that the claims made there can be checked on a CPU:
  * every entry t in (-1, 1) is represented as  sum_{p=1..7} d_p 2^-(8p-2)  with d_1 in [-64, 64] and
    d_2..d_7 in [-128, 127], exactly equal to rint(t 2^54) 2^-54;
  * a product  sum_k a_k b_k  of sliced operands, restricted to the 28 slice pairs with p + q <= 8, differs from
    the product of the rounded operands by less than  K * 6 * 2^-54  (in units of the two scales), and every
    class sum fits an int32 for K <= 16384.
Only tests/ import this file.
"""
import numpy as np

NS = 7
_MASK = np.int64(0x0000808080808080)


def digits(t):
    """t: float64 array in (-1, 1) -> int64 array [7, ...] of the digits d_1..d_7 (same bit trick as digit_bytes)."""
    t = np.asarray(t, dtype=np.float64)
    x = np.rint(t * 2.0 ** 54).astype(np.int64)
    z = ((x + _MASK) ^ _MASK)
    out = np.empty((NS,) + t.shape, dtype=np.int64)
    for k in range(6):                                   # byte k = digit 7 - k
        out[6 - k] = ((z >> np.int64(8 * k)) & np.int64(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64)
    out[0] = z >> np.int64(48)                           # arithmetic shift: d_1
    return out


def undigits(d):
    """exact integer  sum_p d_p 2^(8 (7 - p))  = rint(t 2^54)."""
    acc = np.zeros(d.shape[1:], dtype=object)
    for p in range(NS):
        acc = acc + d[p].astype(object) * (1 << (8 * (6 - p)))
    return acc


def row_scale(a):
    """power of two strictly above max |a| along the last axis (1 for an all-zero row), as row_scale() on the device."""
    mx = np.max(np.abs(a), axis=-1)
    e = np.frexp(mx)[1]
    return np.where(mx > 0, np.ldexp(1.0, e), 1.0)


def sliced_product(A, B, sb=None):
    """A [M, K] (row operand), B [N, K] (column operand; sb: its scales, default one per row of B) -> (V, classes, ...)

    V       float64 [M, N]: what the kernels return, sigma_i tau_j sum_w 2^(4-8w) S_w with the 7 class sums S_w of the
            28 kept slice pairs (combined here exactly and rounded once);
    classes int64 [7, M, N]: the class sums (must fit int32);
    V_full  exact product of the ROUNDED operands as Python integers scaled by 2^108 (object array), for error bounds.
    """
    sa = row_scale(A)
    sb = row_scale(B) if sb is None else np.broadcast_to(np.asarray(sb, dtype=np.float64), (B.shape[0],)).copy()
    da, db = digits(A / sa[:, None]), digits(B / sb[:, None])
    M, N = A.shape[0], B.shape[0]
    classes = np.zeros((NS, M, N), dtype=np.int64)
    for p in range(1, NS + 1):
        for q in range(1, NS + 2 - p):                   # p + q <= 8
            classes[p + q - 2] += da[p - 1] @ db[q - 1].T
    acc = np.zeros((M, N), dtype=object)
    for w in range(2, NS + 2):
        acc = acc + classes[w - 2].astype(object) * (1 << (8 * (8 - w)))      # common denominator 2^60 = 2^(8*8-4)
    V = np.array([[float(acc[i, j]) for j in range(N)] for i in range(M)]) * 2.0 ** -60 * sa[:, None] * sb[None, :]
    xa, xb = undigits(da), undigits(db)
    V_full = xa @ xb.T                                   # exact, scaled by 2^108
    return V, classes, V_full, acc, sa, sb


# ---- block-recursive triangular inverse (the level / pair structure of trtri_i8.cu) --------------------------

def n2_of(NB, nb, pair):
    """row blocks of X22 in pair `pair` at level nb (0: the pair has no second half); trtri_i8.cu n2_of."""
    r = NB - pair * 2 * nb - nb
    return 0 if r < 0 else min(r, nb)


def blockrec_inverse(L, blk):
    """Inverse of the lower-triangular L (order a multiple of blk) by the recursion of trtri_i8.cu:
    level 0 inverts the blk x blk diagonal blocks, level nb = 1, 2, 4, ... sets X21 = -X22 (L21 X11) for every
    pair of neighbouring nb-block halves; the last pair of a level may have a short or empty second half.
    Only the k ranges the kernels visit are used: W = X22 L21 row block `it` meets k blocks 0..it, X21 = -W X11
    column block `nt` meets k blocks nt..nb-1."""
    n = L.shape[0]
    NB = n // blk
    X = np.tril(L).astype(np.float64).copy()
    for j in range(NB):
        s = slice(j * blk, (j + 1) * blk)
        X[s, s] = np.linalg.inv(np.tril(L[s, s]))
    nb = 1
    while nb < NB:
        n_pairs = (NB + 2 * nb - 1) // (2 * nb)
        for pair in range(n_pairs):
            n2 = n2_of(NB, nb, pair)
            if n2 == 0:
                continue
            r0 = pair * 2 * nb
            W = np.zeros((n2 * blk, nb * blk))
            for it in range(n2):                          # W[it] = sum_{k <= it} X22[it, k] L21[k]
                for k in range(it + 1):
                    a = X[(r0 + nb + it) * blk:(r0 + nb + it + 1) * blk, (r0 + nb + k) * blk:(r0 + nb + k + 1) * blk]
                    b = X[(r0 + nb + k) * blk:(r0 + nb + k + 1) * blk, r0 * blk:(r0 + nb) * blk]      # still L21 here
                    W[it * blk:(it + 1) * blk] += a @ b
            out = np.zeros_like(W)
            for nt in range(nb):                          # X21[:, nt] = - sum_{k >= nt} W[:, k] X11[k, nt]
                for k in range(nt, nb):
                    out[:, nt * blk:(nt + 1) * blk] -= W[:, k * blk:(k + 1) * blk] @ X[(r0 + k) * blk:(r0 + k + 1) * blk,
                                                                                        (r0 + nt) * blk:(r0 + nt + 1) * blk]
            X[(r0 + nb) * blk:(r0 + nb + n2) * blk, r0 * blk:(r0 + nb) * blk] = out
        nb *= 2
    return X


# ---- work items of the persistent INT8 posterior kernel (struct Walk in posterior_i8.cu) --------------------

def walk_items(NB, chunk, group, n_tiles):
    """[(tile, [row blocks])] in issue order: groups of `group` tiles, inside a group pair level then tile; pair level
    l takes row-block chunks l and cpt - 1 - l (one chunk when they coincide)."""
    cpt = (NB + chunk - 1) // chunk
    npl = (cpt + 1) // 2
    items = []
    for item in range(n_tiles * npl):
        g = min(item // (group * npl), (n_tiles - 1) // group)
        g_tiles = min(group, n_tiles - g * group)
        within = item - g * group * npl
        tile, level = g * group + within % g_tiles, within // g_tiles
        parts = [level] if 2 * level == cpt - 1 else [level, cpt - 1 - level]
        rbs = [rb for c in parts for rb in range(c * chunk, min(NB, (c + 1) * chunk))]
        items.append((tile, rbs))
    return items
