// Outer-panel trailing update of the blocked Cholesky on the INT8 tensor pipe.
//
// b7_launch_potrf (potrf.cu) factors an outer panel of W = 4 blocks (512 columns) with the latency-bound
// diag / panel chain and then applies it to everything on its right, C[it][nt] -= P[it] P[nt]^T with k = 512:
// 83 % of the flops of the factorisation, 19.6 of its 28.8 ms at N = 4096, S = 32 on FP64 DMMA tiles (80 % of the
// FP64 peak).  Here the same update runs through the error-free slicing of posterior_i8.cu: every row of the
// panel is split once into 7 radix-256 int8 slices with its own power-of-two scale sigma_r (slice_panel_kernel),
// the 28 slice products with p + q <= 8 are exact int32 dot products on tcgen05.mma.kind::i8 (|d e| <= 2^14,
// 7 products x 512 terms per class), and the epilogue rebuilds the fp64 value and subtracts it from C with one
// rounding:  C <- fma(-v, sigma_i sigma_j, C).  The truncated products are below 2^-54 sigma_i sigma_j per term,
// i.e. of the size of the rounding of the fp64 accumulation itself (sigma_r^2 <= 4 K_rr).
//
// trail_i8_kernel: a CTA takes a chunk of consecutive work items (TMEM allocated once per CTA).  An item is one 128 x 64 piece of C
// (tile (it, nt), half h): warp 4 streams the stages (56 KB of P[it] in the 128-row UMMA layout + 28 KB of the
// 64-row half of P[nt]; the slice kernel writes both layouts), warp 5 issues 56 MMAs per stage from one
// elect.sync lane with the A tile held in the collector, warps 0-3 drain the 7 class accumulators once per item,
// free TMEM for the next item and then do the read-modify-write of C (coalesced 32-byte pieces of the tiled
// fp64 layout) while the next item's MMAs run.
#include <math.h>
#include <stdlib.h>

#include "b7_internal.h"
#include "gemm_tile.cuh"
#include "i8_common.cuh"

using b7g::mbar_init; using b7g::mbar_wait; using b7g::mbar_arrive; using b7g::mbar_arrive_expect_tx; using b7g::bulk_g2s;
using b7g::mbar_fence_init; using b7g::smem_u32; using b7g::tile_off; using b7g::elem_off;
using namespace b7i8;

namespace {

constexpr int TM = 128, TN = 64, KB = 64, KC = KB / 16;
constexpr int A_STAGE = NS * TM * KB;      // 57344 B
constexpr int B_STAGE = NS * TN * KB;      // 28672 B
constexpr int STAGE = A_STAGE + B_STAGE;   // 86016 B
constexpr int NSTAGE = 2;
constexpr int T_THREADS = 192;
constexpr int T_SMEM = NSTAGE * STAGE + 1024;

// rows [row0_blk*128, Np) of the panel columns [k0, k0 + n_ks*64): per-row scale and slices in both layouts
//   pA[rb - row0_blk][ks][p][kc][128 rows][16]      pB[(rb - row0_blk) * 2 + half][ks][p][kc][64 rows][16]
__global__ void __launch_bounds__(SLICE_THREADS)
slice_panel_kernel(const double* __restrict__ fac, long long fac_stride, int Np, int kb0, int n_ks, int row0_blk,
                   int8_t* __restrict__ pA, int8_t* __restrict__ pB, long long p_stride, double* __restrict__ sig, int s0) {
  __shared__ double s_red[SLICE_THREADS];
  const int rb = row0_blk + blockIdx.x, s = s0 + blockIdx.y, row = threadIdx.x & 127, q = threadIdx.x >> 7;
  const int KTA = Np / 16, n_kc = n_ks * KC;
  const double* src = fac + (long long)s * fac_stride + tile_off(KTA, rb, kb0 * (128 / 16));
  double mx = 0.0;
  bool bad = false;
  for (int kc = q; kc < n_kc; kc += 4) chunk_max(src + elem_off(row, kc * 16), mx, bad);
  const double sg = row_scale(mx, bad, s_red, row, q);
  const double inv = sg != sg ? 0.0 : 1.0 / sg;
  if (q == 0) sig[(long long)blockIdx.y * Np + rb * TM + row] = sg;   // scratch is indexed by the draw's position in the batch
  int8_t* dA = pA + (long long)blockIdx.y * p_stride + (long long)blockIdx.x * n_ks * A_STAGE;
  int8_t* dB = pB + (long long)blockIdx.y * p_stride + (long long)(blockIdx.x * 2 + (row >> 6)) * n_ks * B_STAGE;
  for (int kc = q; kc < n_kc; kc += 4) {
    uint32_t pk[4][NS];
    chunk_digits(src + elem_off(row, kc * 16), inv, pk);
    const int ks = kc / KC, kcc = kc % KC;
#pragma unroll
    for (int p = 0; p < NS; ++p) {
      const uint4 v = make_uint4(pk[0][p], pk[1][p], pk[2][p], pk[3][p]);
      *reinterpret_cast<uint4*>(dA + (long long)ks * A_STAGE + p * (KC * TM * 16) + kcc * (TM * 16) + row * 16) = v;
      *reinterpret_cast<uint4*>(dB + (long long)ks * B_STAGE + p * (KC * TN * 16) + kcc * (TN * 16) + (row & 63) * 16) = v;
    }
  }
}

struct Item { int s, sr, it, nt, h; bool live; };   // s: draw, sr: its position in the batch

__device__ __forceinline__ Item decode(long long w, int n_it, int n_nt, int it0, int nt0, int s0) {
  const int per = n_it * n_nt * 2;
  Item x;
  const int r = (int)(w % per);
  x.sr = (int)(w / per);
  x.s = s0 + x.sr;
  x.it = it0 + r / (n_nt * 2);
  x.nt = nt0 + (r >> 1) % n_nt;
  x.h = r & 1;
  x.live = x.nt <= x.it;
  return x;
}

__global__ void __launch_bounds__(T_THREADS, 1)
trail_i8_kernel(double* __restrict__ fac, long long fac_stride, int Np, const int8_t* __restrict__ pA, const int8_t* __restrict__ pB,
                long long p_stride, const double* __restrict__ sig, int row0_blk, int n_ks, int it0, int n_it, int nt0, int n_nt,
                int s0, long long n_items, int chunk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t *full = bars, *empty = bars + NSTAGE, *acc_full = bars + 2 * NSTAGE, *acc_empty = bars + 2 * NSTAGE + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // a CTA owns `chunk` consecutive items (they share the row block `it`, so its slices stay in L2) and then exits:
  // the far update runs next to the latency chain of the next panel, which must be able to claim SMs
  const long long w_begin = (long long)blockIdx.x * chunk;
  const long long w_end = w_begin + chunk < n_items ? w_begin + chunk : n_items;
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
    // ---- producer ----
    int slot = 0;
    unsigned phase = 1;
    bool wrapped = false;
    for (long long w = w_begin; w < w_end; ++w) {
      const Item x = decode(w, n_it, n_nt, it0, nt0, s0);
      if (!x.live) continue;
      const int8_t* a = pA + (long long)x.sr * p_stride + (long long)(x.it - row0_blk) * n_ks * A_STAGE;
      const int8_t* b = pB + (long long)x.sr * p_stride + (long long)((x.nt - row0_blk) * 2 + x.h) * n_ks * B_STAGE;
      for (int ks = 0; ks < n_ks; ++ks) {
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
    // ---- MMA issuer ----
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    int slot = 0;
    unsigned phase = 0;
    int done = 0;
    for (long long w = w_begin; w < w_end; ++w) {
      const Item x = decode(w, n_it, n_nt, it0, nt0, s0);
      if (!x.live) continue;
      if (done > 0) { mbar_wait(acc_empty, (unsigned)((done - 1) & 1)); tc_fence_after(); }
      for (int ks = 0; ks < n_ks; ++ks) {
        mbar_wait(full + slot, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint8_t* sa = smem + slot * STAGE;
          const uint64_t da0 = umma_desc(sa, TM * 16, 128), db0 = umma_desc(sa + A_STAGE, TN * 16, 128);
#pragma unroll
          for (int k2 = 0; k2 < KB / 32; ++k2) {
#pragma unroll
            for (int p = 1; p <= NS; ++p) {
              const uint64_t da = da0 + (uint64_t)(((p - 1) * (KC * TM * 16) + k2 * (2 * TM * 16)) >> 4);
              const uint32_t acc = (ks == 0 && k2 == 0 && p == 1) ? 0u : 1u;
              const int nq = NS + 1 - p;
#pragma unroll
              for (int q = 1; q <= nq; ++q) {
                const uint64_t db = db0 + (uint64_t)(((q - 1) * (KC * TN * 16) + k2 * (2 * TN * 16)) >> 4);
                const uint32_t dcol = tmem + (uint32_t)((p + q - 2) * TN);
                if (nq == 1) umma_i8<0>(dcol, da, db, idesc, acc);
                else if (q == 1) umma_i8<1>(dcol, da, db, idesc, acc);
                else if (q == nq) umma_i8<3>(dcol, da, db, idesc, acc);
                else umma_i8<2>(dcol, da, db, idesc, acc);
              }
            }
          }
          umma_commit(empty + slot);
          if (ks == n_ks - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++slot == NSTAGE) { slot = 0; phase ^= 1u; }
      }
      ++done;
    }
  } else {
    // ---- epilogue warps 0-3: thread = row of the C tile (TMEM lane), 64 columns ----
    int done = 0;
    for (long long w = w_begin; w < w_end; ++w) {
      const Item x = decode(w, n_it, n_nt, it0, nt0, s0);
      if (!x.live) continue;
      mbar_wait(acc_full, (unsigned)(done & 1));
      tc_fence_after();
      double v[TN];
#pragma unroll
      for (int c = 0; c < TN; ++c) v[c] = 0.0;
#pragma unroll
      for (int wc = 2; wc <= NS + 1; ++wc) {
        const double wt = ldexp(1.0, 4 - 8 * wc);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t dv[32];
          tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((wc - 2) * TN + hh * 32), dv);
#pragma unroll
          for (int c = 0; c < 32; ++c) v[hh * 32 + c] = fma((double)(int)dv[c], wt, v[hh * 32 + c]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);     // TMEM drained: the next item's MMAs may start
      ++done;
      const double* sg = sig + (long long)x.sr * Np;
      const double si = sg[x.it * TM + tid];
      const double* sj = sg + x.nt * TM + x.h * TN;
      double* C = fac + (long long)x.s * fac_stride + tile_off(Np / 16, x.it, x.nt * (128 / 16) + x.h * (TN / 16));
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        double4* cp = reinterpret_cast<double4*>(C + elem_off(tid, g * 4));
        double4 c = *cp;
        c.x = fma(-v[g * 4 + 0], si * sj[g * 4 + 0], c.x);
        c.y = fma(-v[g * 4 + 1], si * sj[g * 4 + 1], c.y);
        c.z = fma(-v[g * 4 + 2], si * sj[g * 4 + 2], c.z);
        c.w = fma(-v[g * 4 + 3], si * sj[g * 4 + 3], c.w);
        *cp = c;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

bool g_attr[16] = {false};

}  // namespace

size_t b7_i8_panel_bytes(int Np, int W) { return (size_t)Np * (size_t)(W * 128) * NS; }

int b7_i8_panel_slice(b7_ctx* ctx, cudaStream_t st, const double* fac, int Np, int kb0, int kb1, int row0_blk, int8_t* pA, int8_t* pB,
                      size_t p_stride, double* sig, int s0, int count) {
  const int NB = Np / 128;
  if (row0_blk >= NB) return 0;
  slice_panel_kernel<<<dim3(NB - row0_blk, count), SLICE_THREADS, 0, st>>>(fac, (long long)Np * Np, Np, kb0, (kb1 - kb0) * 2, row0_blk, pA, pB,
                                                                (long long)p_stride, sig, s0);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

int b7_i8_trail(b7_ctx* ctx, cudaStream_t st, double* fac, int Np, const int8_t* pA, const int8_t* pB, size_t p_stride, const double* sig,
                int kb0, int kb1, int row0_blk, int it0, int n_it, int nt0, int n_nt, int s0, int count) {
  if (!g_attr[ctx->device & 15]) {
    B7_CUDA(cudaFuncSetAttribute(trail_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
    g_attr[ctx->device & 15] = true;
  }
  if (n_it <= 0 || n_nt <= 0 || count <= 0) return 0;
  const long long n_items = (long long)count * n_it * n_nt * 2;
  static const int chunk_env = getenv("B7_TRAIL_CHUNK") ? atoi(getenv("B7_TRAIL_CHUNK")) : 0;
  const int chunk = chunk_env > 0 ? chunk_env : 8;
  const int grid = (int)((n_items + chunk - 1) / chunk);
  trail_i8_kernel<<<grid, T_THREADS, T_SMEM, st>>>(fac, (long long)Np * Np, Np, pA, pB, (long long)p_stride, sig, row0_blk, (kb1 - kb0) * 2,
                                                   it0, n_it, nt0, n_nt, s0, n_items, chunk);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}
