// Inversion of the Cholesky factors on the INT8 tensor pipe.
//
// The sweep of potrf.cu (b7_launch_trtri) multiplies by columns that were computed one step earlier, so its operands
// cannot be sliced ahead of time.  The block-recursive form has only static operands per level:
//     L = | L11  0  |      L^-1 = |         X11            0  |
//         | L21 L22 |             | -X22 (L21 X11)        X22 |
// with X11, X22 the inverses from the previous level (level 0: the 128 x 128 diagonal inverses the factorisation
// already produced).  Level nb = 1, 2, 4, ... pairs neighbouring diagonal blocks of nb x 128 rows; per level
//     W   = X22 L21      (X22 lower triangular: row block `it` only meets k blocks 0 .. it)
//     X21 = - W X11      (X11 lower triangular: column block `nt` only meets k blocks nt .. nb-1)
// for all pairs and all draws at once, in place (X21 overwrites L21).  Same N^3/3 flops as the sweep, 3/4 of them in
// the last level.  Every product runs as 28 exact int8 slice products (see posterior_i8.cu / potrf_i8.cu): the row
// operand is sliced with one power-of-two scale per row, the column operand with one per column, and the epilogue
// writes  sign * v * sigma_row * sigma_col  (one rounding per entry).  A non-power-of-two number of blocks leaves the
// last pair of a level with a short (or no) X22.
#include <math.h>
#include <stdlib.h>

#include "b7_internal.h"
#include "gemm_tile.cuh"
#include "i8_common.cuh"

using b7g::mbar_init; using b7g::mbar_wait; using b7g::mbar_arrive; using b7g::mbar_arrive_expect_tx; using b7g::bulk_g2s;
using b7g::mbar_fence_init; using b7g::smem_u32; using b7g::tile_off; using b7g::elem_off;
using namespace b7i8;
using namespace b7i8::gemm;

namespace {

// row blocks of X22 in pair `pair` at level nb (0: the pair has no second half)
__host__ __device__ __forceinline__ int n2_of(int NB, int nb, int pair) {
  const int r = NB - pair * 2 * nb - nb;
  return r < 0 ? 0 : (r < nb ? r : nb);
}

struct Src {            // a block matrix in the tiled fp64 layout, addressed per pair
  const double* base;   // first draw of the batch
  long long draw;       // doubles between draws
  int kta;              // 16-column tiles per row block
  int rb0, pair_rb;     // row block    = rb0 + pair * pair_rb + i
  int cb0, pair_cb;     // column block = cb0 + pair * pair_cb + j
};

// rows -> A-layout slices  sA[pair * nb + it][ks][p][kc][128][16]  and sig[(pair * nb + it) * 128 + row]
// k blocks of row block `it`: 0 .. it when the operand is lower triangular, else 0 .. nb-1
__global__ void __launch_bounds__(SLICE_THREADS)
slice_rows_kernel(Src m, int NB, int nb, int tri, int8_t* __restrict__ sA, long long sA_draw, double* __restrict__ sig, long long sig_draw) {
  __shared__ double s_red[SLICE_THREADS];
  const int it = blockIdx.x, pair = blockIdx.y, z = blockIdx.z, row = threadIdx.x & 127, q = threadIdx.x >> 7;
  if (it >= n2_of(NB, nb, pair)) return;
  const int n_kc = (tri ? it + 1 : nb) * 8, KS = 2 * nb;
  const double* src = m.base + (long long)z * m.draw + tile_off(m.kta, m.rb0 + pair * m.pair_rb + it, (m.cb0 + pair * m.pair_cb) * 8);
  double mx = 0.0;
  bool bad = false;
  for (int kc = q; kc < n_kc; kc += 4) chunk_max(src + elem_off(row, kc * 16), mx, bad);
  const double sg = row_scale(mx, bad, s_red, row, q);
  const double inv = sg != sg ? 0.0 : 1.0 / sg;
  if (q == 0) sig[(long long)z * sig_draw + (pair * nb + it) * TM + row] = sg;
  int8_t* dst = sA + (long long)z * sA_draw + (long long)(pair * nb + it) * KS * A_STAGE;
  for (int kc = q; kc < n_kc; kc += 4) {
    uint32_t pk[4][NS];
    chunk_digits(src + elem_off(row, kc * 16), inv, pk);
    const int ks = kc / KC, kcc = kc % KC;
#pragma unroll
    for (int p = 0; p < NS; ++p)
      *reinterpret_cast<uint4*>(dst + (long long)ks * A_STAGE + p * (KC * TM * 16) + kcc * (TM * 16) + row * 16) =
          make_uint4(pk[0][p], pk[1][p], pk[2][p], pk[3][p]);
  }
}

// columns -> B-layout slices  sB[(pair * nb + nt) * 2 + half][ks][p][kc][64][16]  and sig[(pair * nb + nt) * 128 + col]
// k (row) blocks of column block `nt`: nt .. nb-1 when the operand is lower triangular (X11), else 0 .. n2-1 (L21)
__global__ void __launch_bounds__(SLICE_THREADS)
slice_cols_kernel(Src m, int NB, int nb, int tri, int8_t* __restrict__ sB, long long sB_draw, double* __restrict__ sig, long long sig_draw) {
  __shared__ double s_red[SLICE_THREADS];
  const int nt = blockIdx.x, pair = blockIdx.y, z = blockIdx.z, n = threadIdx.x & 127, q = threadIdx.x >> 7;
  const int n2 = n2_of(NB, nb, pair);
  if (n2 == 0) return;
  const int kc0 = (tri ? nt : 0) * 8, kc1 = (tri ? nb : n2) * 8, KS = 2 * nb;
  const double* base = m.base + (long long)z * m.draw;
  const int ct = (m.cb0 + pair * m.pair_cb + nt) * 8 + (n >> 4);
  const int coff = ((n & 15) >> 2) * (TM * 4) + (n & 3);            // elem_off(k, n & 15) = coff + 4 k
  double mx = 0.0;
  bool bad = false;
  for (int kc = kc0 + q; kc < kc1; kc += 4) {
    const double* t = base + tile_off(m.kta, m.rb0 + pair * m.pair_rb + (kc >> 3), ct) + coff + 4 * 16 * (kc & 7);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const double x = t[4 * k];
      bad |= !isfinite(x);
      mx = fmax(mx, fabs(x));
    }
  }
  const double sg = row_scale(mx, bad, s_red, n, q);
  const double inv = sg != sg ? 0.0 : 1.0 / sg;
  if (q == 0) sig[(long long)z * sig_draw + (pair * nb + nt) * TM + n] = sg;
  int8_t* dst = sB + (long long)z * sB_draw + (long long)((pair * nb + nt) * 2 + (n >> 6)) * KS * B_STAGE + (n & 63) * 16;
  for (int kc = kc0 + q; kc < kc1; kc += 4) {
    const double* t = base + tile_off(m.kta, m.rb0 + pair * m.pair_rb + (kc >> 3), ct) + coff + 4 * 16 * (kc & 7);
    uint32_t pk[4][NS];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      unsigned long long zz[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) zz[i] = digit_bytes(t[4 * (g * 4 + i)] * inv);
      pack4(zz, pk[g]);
    }
    const int ks = kc / KC, kcc = kc % KC;
#pragma unroll
    for (int p = 0; p < NS; ++p)
      *reinterpret_cast<uint4*>(dst + (long long)ks * B_STAGE + p * (KC * TN * 16) + kcc * (TN * 16)) =
          make_uint4(pk[0][p], pk[1][p], pk[2][p], pk[3][p]);
  }
}

struct Out {
  double* base;         // first draw of the batch
  long long draw;
  int kta, rb0, pair_rb, cb0, pair_cb;
};

struct Item { int z, pair, it, nt, h, ks0, ks1; bool live; };

__device__ __forceinline__ Item decode(long long w, int NB, int nb, int n_pairs, int mode) {
  const int per_pair = nb * nb * 2, per_draw = per_pair * n_pairs;
  Item x;
  x.z = (int)(w / per_draw);
  int r = (int)(w % per_draw);
  x.pair = r / per_pair;
  r -= x.pair * per_pair;
  x.it = r / (nb * 2);
  x.nt = (r >> 1) % nb;
  x.h = r & 1;
  x.live = x.it < n2_of(NB, nb, x.pair);
  x.ks0 = mode == 0 ? 0 : 2 * x.nt;
  x.ks1 = mode == 0 ? 2 * (x.it + 1) : 2 * nb;
  return x;
}

// out tile (it, nt, half) = sign * sum_{ks0 <= ks < ks1} A[it][ks] B[nt, half][ks]^T * sigA * sigB
__global__ void __launch_bounds__(THREADS, 1)
gemm_i8_kernel(const int8_t* __restrict__ sA, long long sA_draw, const int8_t* __restrict__ sB, long long sB_draw,
               const double* __restrict__ sigA, const double* __restrict__ sigB, long long sig_draw, Out o, int NB, int nb, int n_pairs,
               int mode, double sign, long long n_items, int chunk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t *full = bars, *empty = bars + NSTAGE, *acc_full = bars + 2 * NSTAGE, *acc_empty = bars + 2 * NSTAGE + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long w_begin = (long long)blockIdx.x * chunk;
  const long long w_end = w_begin + chunk < n_items ? w_begin + chunk : n_items;
  const int KS = 2 * nb;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    mbar_fence_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    int slot = 0;
    unsigned phase = 1;
    bool wrapped = false;
    for (long long w = w_begin; w < w_end; ++w) {
      const Item x = decode(w, NB, nb, n_pairs, mode);
      if (!x.live) continue;
      const int8_t* a = sA + (long long)x.z * sA_draw + (long long)(x.pair * nb + x.it) * KS * A_STAGE;
      const int8_t* b = sB + (long long)x.z * sB_draw + (long long)((x.pair * nb + x.nt) * 2 + x.h) * KS * B_STAGE;
      for (int ks = x.ks0; ks < x.ks1; ++ks) {
        if (wrapped) mbar_wait(empty + slot, phase);
        if (elect_one()) {
          uint8_t* st = smem + slot * STAGE;
          mbar_arrive_expect_tx(full + slot, STAGE);
          bulk_g2s(st, a + (long long)ks * A_STAGE, A_STAGE, full + slot);
          bulk_g2s(st + A_STAGE, b + (long long)ks * B_STAGE, B_STAGE, full + slot);
        }
        __syncwarp();
        if (++slot == NSTAGE) { slot = 0; phase ^= 1u; wrapped = true; }
      }
    }
  } else if (warp == 5) {
    const uint32_t idesc = idesc_128x64();
    int slot = 0, done = 0;
    unsigned phase = 0;
    for (long long w = w_begin; w < w_end; ++w) {
      const Item x = decode(w, NB, nb, n_pairs, mode);
      if (!x.live) continue;
      if (done > 0) { mbar_wait(acc_empty, (unsigned)((done - 1) & 1)); tc_fence_after(); }
      for (int ks = x.ks0; ks < x.ks1; ++ks) {
        mbar_wait(full + slot, phase);
        tc_fence_after();
        if (elect_one()) {
          mma_stage(smem + slot * STAGE, tmem, idesc, ks == x.ks0);
          umma_commit(empty + slot);
          if (ks == x.ks1 - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++slot == NSTAGE) { slot = 0; phase ^= 1u; }
      }
      ++done;
    }
  } else {
    int done = 0;
    for (long long w = w_begin; w < w_end; ++w) {
      const Item x = decode(w, NB, nb, n_pairs, mode);
      if (!x.live) continue;
      mbar_wait(acc_full, (unsigned)(done & 1));
      tc_fence_after();
      double v[TN];
      drain_classes(tmem, warp, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      ++done;
      const double si = sign * sigA[(long long)x.z * sig_draw + (x.pair * nb + x.it) * TM + tid];
      const double* sj = sigB + (long long)x.z * sig_draw + (x.pair * nb + x.nt) * TM + x.h * TN;
      double* C = o.base + (long long)x.z * o.draw +
                  tile_off(o.kta, o.rb0 + x.pair * o.pair_rb + x.it, (o.cb0 + x.pair * o.pair_cb + x.nt) * 8 + x.h * (TN / 16));
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        double4 c;
        c.x = v[g * 4 + 0] * (si * sj[g * 4 + 0]);
        c.y = v[g * 4 + 1] * (si * sj[g * 4 + 1]);
        c.z = v[g * 4 + 2] * (si * sj[g * 4 + 2]);
        c.w = v[g * 4 + 3] * (si * sj[g * 4 + 3]);
        *reinterpret_cast<double4*>(C + elem_off(tid, g * 4)) = c;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

bool g_attr[16] = {false};

int launch_gemm(b7_ctx* ctx, const int8_t* sA, long long sA_draw, const int8_t* sB, long long sB_draw, const double* sigA,
                const double* sigB, long long sig_draw, Out o, int NB, int nb, int n_pairs, int mode, double sign, int count) {
  const long long n_items = (long long)count * n_pairs * nb * nb * 2;
  static const int chunk_env = getenv("B7_TRTRI_CHUNK") ? atoi(getenv("B7_TRTRI_CHUNK")) : 0;
  const int chunk = chunk_env > 0 ? chunk_env : 4;   // fewer draws in flight: their slices stay in L2 (12.9 vs 13.5 ms at chunk 8)
  const int grid = (int)((n_items + chunk - 1) / chunk);
  gemm_i8_kernel<<<grid, THREADS, SMEM, ctx->stream>>>(sA, sA_draw, sB, sB_draw, sigA, sigB, sig_draw, o, NB, nb, n_pairs, mode, sign, n_items,
                                                       chunk);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// L -> L^-1 in place for draws [s0, s0 + count); the diagonal tiles must already hold the 128 x 128 inverses
static int trtri_i8_group(b7_gp* gp, int s0, int count, size_t rows_max, size_t w_max);

int b7_launch_trtri_i8(b7_gp* gp, int s0, int count) {
  b7_ctx* ctx = gp->ctx;
  const int NB = gp->NB;
  if (NB < 2) return 0;
  if (!g_attr[ctx->device & 15]) {
    B7_CUDA(cudaFuncSetAttribute(gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    g_attr[ctx->device & 15] = true;
  }
  // scratch per draw, sized for the level that needs most: slices (pairs * nb row blocks x 2 nb stages), W, scales
  size_t rows_max = 0, w_max = 0;
  for (int nb = 1; nb < NB; nb *= 2) {
    const size_t pairs = (size_t)(NB + 2 * nb - 1) / (2 * nb);
    rows_max = rows_max > pairs * nb * 2 * nb ? rows_max : pairs * nb * 2 * nb;
    w_max = w_max > pairs * nb * nb ? w_max : pairs * nb * nb;
  }
  // draws are processed in groups so that the scratch stays below ~6 GB whatever N and S are
  const size_t per_draw = rows_max * (A_STAGE + 2 * B_STAGE) + w_max * 128 * 128 * sizeof(double);
  const size_t budget = (size_t)6 << 30;
  int group = (int)(budget / per_draw);
  group = group < 1 ? 1 : (group > count ? count : group);
  for (int g0 = 0; g0 < count; g0 += group)
    B7_CHECK(trtri_i8_group(gp, s0 + g0, count - g0 < group ? count - g0 : group, rows_max, w_max));
  return 0;
}

static int trtri_i8_group(b7_gp* gp, int s0, int count, size_t rows_max, size_t w_max) {
  b7_ctx* ctx = gp->ctx;
  const int Np = gp->Np, NB = gp->NB;
  const size_t sA_draw = rows_max * A_STAGE, sB_draw = rows_max * 2 * B_STAGE, w_draw = w_max * 128 * 128, sig_draw = (size_t)Np;
  int8_t *sA = nullptr, *sB = nullptr;
  double *W = nullptr, *sigA = nullptr, *sigB = nullptr;
  B7_CHECK(b7_pool_alloc(ctx, (void**)&sA, sA_draw * count));
  B7_CHECK(b7_pool_alloc(ctx, (void**)&sB, sB_draw * count));
  B7_CHECK(b7_pool_alloc(ctx, (void**)&W, w_draw * count * sizeof(double)));
  B7_CHECK(b7_pool_alloc(ctx, (void**)&sigA, sig_draw * count * sizeof(double)));
  B7_CHECK(b7_pool_alloc(ctx, (void**)&sigB, sig_draw * count * sizeof(double)));
  const long long fs = (long long)Np * Np;
  double* fac0 = gp->fac + (long long)s0 * fs;
  cudaStream_t st = ctx->stream;
  for (int nb = 1; nb < NB; nb *= 2) {
    const int n_pairs = (NB + 2 * nb - 1) / (2 * nb);
    const dim3 grid(nb, n_pairs, count);
    const Src x22{fac0, fs, Np / 16, nb, 2 * nb, nb, 2 * nb};
    const Src l21{fac0, fs, Np / 16, nb, 2 * nb, 0, 2 * nb};
    const Src x11{fac0, fs, Np / 16, 0, 2 * nb, 0, 2 * nb};
    const Src wsrc{W, (long long)w_draw, nb * 8, 0, nb, 0, 0};
    // W = X22 L21
    slice_rows_kernel<<<grid, SLICE_THREADS, 0, st>>>(x22, NB, nb, 1, sA, (long long)sA_draw, sigA, (long long)sig_draw);
    slice_cols_kernel<<<grid, SLICE_THREADS, 0, st>>>(l21, NB, nb, 0, sB, (long long)sB_draw, sigB, (long long)sig_draw);
    b7_count(ctx, 2);
    B7_CHECK(launch_gemm(ctx, sA, (long long)sA_draw, sB, (long long)sB_draw, sigA, sigB, (long long)sig_draw,
                         Out{W, (long long)w_draw, nb * 8, 0, nb, 0, 0}, NB, nb, n_pairs, 0, 1.0, count));
    // X21 = - W X11
    slice_rows_kernel<<<grid, SLICE_THREADS, 0, st>>>(wsrc, NB, nb, 0, sA, (long long)sA_draw, sigA, (long long)sig_draw);
    slice_cols_kernel<<<grid, SLICE_THREADS, 0, st>>>(x11, NB, nb, 1, sB, (long long)sB_draw, sigB, (long long)sig_draw);
    b7_count(ctx, 2);
    B7_CHECK(launch_gemm(ctx, sA, (long long)sA_draw, sB, (long long)sB_draw, sigA, sigB, (long long)sig_draw,
                         Out{fac0, fs, Np / 16, nb, 2 * nb, 0, 2 * nb}, NB, nb, n_pairs, 1, -1.0, count));
  }
  b7_pool_free(ctx, sA);
  b7_pool_free(ctx, sB);
  b7_pool_free(ctx, W);
  b7_pool_free(ctx, sigA);
  b7_pool_free(ctx, sigB);
  B7_CUDA(cudaGetLastError());
  return 0;
}
