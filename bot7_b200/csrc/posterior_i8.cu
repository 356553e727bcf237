// Posterior over a candidate panel on the INT8 tensor pipe (the default path; B7_POSTERIOR_I8=0 or
// b7_set_posterior_path(B7_PATH_FP64_DMMA) select posterior.cu).
//
// Same mathematics as posterior.cu (V = L^-1 K*^T, var = sf2 - colsumsq(V), mean = m + V^T beta) and the
// same fp64-level accuracy, but the N^2 product per candidate runs on tcgen05.mma.kind::i8 (measured
// 4.2 POPS on B200, tools/i8_mma_probe.cu, against 37 TFLOP/s for the FP64 DMMA pipe) through an
// error-free slicing of both operands (the "Ozaki scheme"):
//     a_ik = sigma_i * sum_{p=1..7} d_p(i,k) 2^-(8p-2),   d_p integers in [-128, 127]  (54 bits + sign per entry)
//     b_ck = tau     * sum_{q=1..7} e_q(c,k) 2^-(8q-2)
//     v_ic = sigma_i tau sum_{w=2..8} 2^(4-8w) sum_{p+q=w} sum_k d_p(i,k) e_q(c,k)
// Every inner sum is an exact int32 dot product (|d e| <= 2^14, a class holds at most 7 products, so
// k <= 16384 terms fit), the 28 slice pairs with p + q <= 8 are kept (the dropped ones are below 2^-54 of
// sigma_i tau per term), and the 7 weight classes are accumulated in 7 x 64 = 448 TMEM columns and combined
// in fp64 in the epilogue.  Measured against the fp64 path: variance within 5e-14 sf2, mean within 5e-11 at
// N = 4096 (profiles/i8_accuracy_r01.json; tests assert 1e-12 sf2 / 1e-10).
//
// A work item is a tile of 64 candidates x a few row blocks of L^-1 (128 rows = TMEM lanes); a persistent grid
// (one CTA per SM) walks the items.  Warp roles: warp 4 streams the slices (already in the UMMA canonical K-major
// layout in HBM, so a stage is two bulk copies), warp 5 issues the MMAs of a stage (both from one elect.sync lane
// of the converged warp), warps 0-3 drain the accumulators once per row block (tcgen05.ld), rebuild v in fp64,
// reduce v^2 and v*beta over the 128 rows and leave one partial per candidate and row block;
// posterior_i8_finish_kernel adds the partials in row-block order.  The MMAs are issued slice-of-L^-1-major so
// that the 4 KB A operand stays in the tensor core's collector while it meets its 8 - p partner slices of K*
// (tcgen05.mma ... collector::a::fill / use / lastuse).
#include <math.h>
#include <stdlib.h>

#include "b7_internal.h"
#include "exp_neg.cuh"
#include "gemm_tile.cuh"
#include "i8_common.cuh"

using b7g::mbar_init; using b7g::mbar_wait; using b7g::mbar_arrive; using b7g::mbar_arrive_expect_tx; using b7g::bulk_g2s;
using b7g::mbar_fence_init; using b7g::smem_u32;
using namespace b7i8;

namespace {

constexpr int TM = 128, TN = 64, KB = 64;  // L^-1 rows per block, candidates per tile, k bytes per stage
constexpr int KC = KB / 16;                // 16-byte k chunks per stage
constexpr int A_STAGE = NS * TM * KB;      // 57344 B
constexpr int B_STAGE = NS * TN * KB;      // 28672 B
constexpr int STAGE = A_STAGE + B_STAGE;   // 86016 B
constexpr int NSTAGE = 2;                  // (32-byte stages x 5 were measured slower: 95.6 vs 100 TFLOP/s equivalent)
static_assert(STAGE % 1024 == 0 && KB % 32 == 0 && 64 % KB == 0, "stage geometry");
static_assert(KB == b7i8::gemm::KB && TM == b7i8::gemm::TM && TN == b7i8::gemm::TN && NSTAGE == b7i8::gemm::NSTAGE,
              "posterior_i8_kernel issues its MMAs through the shared stage of i8_common.cuh");
constexpr int I8_THREADS = 192;
constexpr int RED_BYTES = 2 * 4 * TN * 2 * 8;           // [rb parity][warp][candidate][sum v^2, sum v beta]
constexpr int I8_SMEM = NSTAGE * STAGE + RED_BYTES + 1024;   // stages + reduction scratch + barriers

// ---- slicing -------------------------------------------------------------------------------------------

// L^-1 (tiled fp64, lower) -> per-row power-of-two scale sigma and the slice array
// facS[rb][ks][p][kc][row][16]  (ks = KB-column stage, kc = 16-column chunk inside it)
__global__ void __launch_bounds__(SLICE_THREADS)
slice_factor_kernel(const double* __restrict__ fac, long long fac_stride, int Np, int8_t* __restrict__ facS,
                    long long facS_stride, double* __restrict__ sigma, int s0) {
  __shared__ double s_red[SLICE_THREADS];
  const int rb = blockIdx.x, s = s0 + blockIdx.y, row = threadIdx.x & 127, q = threadIdx.x >> 7;
  const int KTA = Np / 16, KS_ALL = Np / KB, n_kc = (rb + 1) * (TM / 16);
  const double* src = fac + (long long)s * fac_stride + b7g::tile_off(KTA, rb, 0);
  double mx = 0.0;
  bool bad = false;
  for (int kc = q; kc < n_kc; kc += 4) chunk_max(src + b7g::elem_off(row, kc * 16), mx, bad);
  // a failed factorisation (NaN / inf in the row) must poison the results like it does on the fp64 path
  const double sg = row_scale(mx, bad, s_red, row, q);
  const double inv = sg != sg ? 0.0 : 1.0 / sg;
  if (q == 0) sigma[(long long)s * Np + rb * TM + row] = sg;
  int8_t* dst = facS + (long long)s * facS_stride + (long long)rb * KS_ALL * A_STAGE;
  for (int kc = q; kc < n_kc; kc += 4) {
    uint32_t pk[4][NS];
    chunk_digits(src + b7g::elem_off(row, kc * 16), inv, pk);
    const int ks = kc / KC, kcc = kc % KC;
#pragma unroll
    for (int p = 0; p < NS; ++p)
      *reinterpret_cast<uint4*>(dst + (long long)ks * A_STAGE + p * (KC * TM * 16) + kcc * (TM * 16) + row * 16) =
          make_uint4(pk[0][p], pk[1][p], pk[2][p], pk[3][p]);
  }
}

// K(X*, X) evaluated and sliced in one pass: ksS[ct][ks][q][kc][cand 64][16]; one block per (candidate tile, 64 columns)
// 128 threads, two 16-column chunks each (21.6 -> 16.9 ms per step against 256 threads x one chunk, together with
// folding sf2 / tau * 2^54 into one factor).
template <int DT, int KERNEL>
__global__ void __launch_bounds__(128)
cov_slices_kernel(const double* __restrict__ A, long long rows, int d, const double* __restrict__ Xt, int N, int Np,
                  const double* __restrict__ par, double inv_tau, int8_t* __restrict__ ksS) {
  __shared__ double s_x[DT][64];
  __shared__ double s_w[DT];
  __shared__ double s_tab[64];
  if (threadIdx.x < 64) s_tab[threadIdx.x] = c_exp_tab[threadIdx.x];
  const int ct = blockIdx.x, kb64 = blockIdx.y, KS_ALL = Np / KB;
  const int c = threadIdx.x & 63;                                 // warp = 32 consecutive candidates, one k chunk
  for (int e = threadIdx.x; e < DT * 64; e += 128) {
    const int i = e / 64, k = kb64 * 64 + e % 64;
    s_x[i][e % 64] = (i < d && k < N) ? Xt[(long long)i * Np + k] : 0.0;
  }
  if (threadIdx.x < DT) s_w[threadIdx.x] = threadIdx.x < d ? par[threadIdx.x] : 0.0;
  __syncthreads();
  const long long row = (long long)ct * TN + c;
  double a[DT];
#pragma unroll
  for (int i = 0; i < DT; ++i) a[i] = (row < rows && i < d) ? A[row * d + i] : 0.0;
  // sf2 / tau * 2^54: tau is a power of two, so scaling sf2 first rounds exactly like scaling the product afterwards
  const double sf2s = par[B7_MAX_DIMS] * inv_tau * 18014398509481984.0;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
  const int kc = (threadIdx.x >> 6) + 2 * half;
  uint32_t pk[4][NS];
  unsigned long long z[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int kk = kc * 16 + i, k = kb64 * 64 + kk;
    double val = 0.0;
    if (row < rows && k < N) {
      double r2 = 0.0;
#pragma unroll
      for (int j = 0; j < DT; ++j) {
        const double t = (a[j] - s_x[j][kk]) * s_w[j];
        r2 = fma(t, t, r2);
      }
      if (KERNEL == B7_KERNEL_ARDSE) {
        val = sf2s * exp_neg(-0.5 * r2, s_tab);
      } else {
        const double rr = sqrt(r2), s5r = 2.23606797749978969641 * rr;
        val = sf2s * ((1.0 + s5r + (5.0 / 3.0) * r2) * exp_neg(-s5r, s_tab));
      }
    }
    z[i & 3] = digit_bytes_scaled(val);
    if ((i & 3) == 3) pack4(z, pk[i >> 2]);
  }
  const int gkc = kb64 * 4 + kc, ks = gkc / KC, kcc = gkc % KC;
  int8_t* dst = ksS + ((long long)ct * KS_ALL + ks) * B_STAGE + kcc * (TN * 16) + c * 16;
#pragma unroll
  for (int p = 0; p < NS; ++p) *reinterpret_cast<uint4*>(dst + p * (KC * TN * 16)) = make_uint4(pk[0][p], pk[1][p], pk[2][p], pk[3][p]);
  }
}

// sum over the 32 lanes of x[j] for every j, result for column j lands in lane j (transpose-reduce butterfly)
__device__ __forceinline__ double lane_transpose_sum(double (&x)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const double send = up ? x[j] : x[j + s];
      const double keep = up ? x[j + s] : x[j];
      x[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return x[0];
}

// Enumeration of the work items of posterior_i8_kernel (same on every warp role).  Tiles are taken in groups of
// `group`; inside a group the order is pair level, then tile.  Pair level l = row-block chunks l and cpt-1-l.
struct Walk {
  int NB, chunk, group, n_tiles;
  __device__ __forceinline__ int cpt() const { return (NB + chunk - 1) / chunk; }          // chunks per tile
  __device__ __forceinline__ int npl() const { return (cpt() + 1) / 2; }                   // pair levels per tile
  __device__ __forceinline__ int n_items() const { return n_tiles * npl(); }
  __device__ __forceinline__ int grp(int item) const {
    const int g = item / (group * npl()), last = (n_tiles - 1) / group;
    return g < last ? g : last;
  }
  __device__ __forceinline__ int g_tiles(int g) const { return n_tiles - g * group < group ? n_tiles - g * group : group; }
  __device__ __forceinline__ int tile(int item) const {
    const int g = grp(item);
    return g * group + (item - g * group * npl()) % g_tiles(g);
  }
  __device__ __forceinline__ int level(int item) const {
    const int g = grp(item);
    return (item - g * group * npl()) / g_tiles(g);
  }
  __device__ __forceinline__ int parts(int item) const { return 2 * level(item) == cpt() - 1 ? 1 : 2; }
  __device__ __forceinline__ int rb_begin(int item, int part) const { return (part == 0 ? level(item) : cpt() - 1 - level(item)) * chunk; }
  __device__ __forceinline__ int rb_end(int item, int part) const {
    const int e = rb_begin(item, part) + chunk;
    return e < NB ? e : NB;
  }
};

// ---- the kernel ----------------------------------------------------------------------------------------

__global__ void __launch_bounds__(I8_THREADS, 1)
posterior_i8_kernel(const int8_t* __restrict__ facS, const double* __restrict__ sigma, const double* __restrict__ beta,
                    int Np, int NB, const int8_t* __restrict__ ksS, double tau, int chunk, int group, int n_tiles,
                    double* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem[];
  double* red = reinterpret_cast<double*>(smem + NSTAGE * STAGE);             // [2 parity][4 warps][64 cols][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE + RED_BYTES);
  uint64_t *full = bars, *empty = bars + NSTAGE, *acc_full = bars + 2 * NSTAGE, *acc_empty = bars + 2 * NSTAGE + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS_ALL = Np / KB;
  // Work items of equal weight: a 64-candidate tile x a pair of row-block chunks (chunk l and chunk cpt-1-l: the
  // triangular L^-1 makes their stage counts add up to the same number for every l).  Each item leaves one partial
  // (sum v^2, sum v beta) per candidate and row block; posterior_i8_finish_kernel adds them in row-block order.
  // Items are ordered in groups of `group` tiles and dealt round-robin to a persistent grid (one CTA per SM, the
  // TMA / MMA / epilogue pipeline never drains between items): only one or two groups are in flight, so their K*
  // slices (1.8 MB per tile at N = 4096, re-read once per row block) stay in L2 next to the L^-1 slices.
  const Walk wk{NB, chunk, group, n_tiles};
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
    // ---- producer: stream (rb, ks) stages.  The whole warp walks the loop and one elected lane issues the
    // copies: under `if (lane == 0)` the compiler wraps every uniform-datapath instruction (UBLKCP, UTCIMMA) in
    // an ELECT / BRA.U.ANY serialisation loop, which made the MMA issue the bottleneck (61 clk per MMA). ----
    int slot = 0;
    unsigned phase = 1;                    // parity of the *previous* use of the slot; first round needs no wait
    bool wrapped = false;
    for (int item = blockIdx.x; item < wk.n_items(); item += gridDim.x) {
      const int tile = wk.tile(item);
      const int8_t* gB = ksS + (long long)tile * KS_ALL * B_STAGE;
      for (int part = 0; part < wk.parts(item); ++part)
        for (int rb = wk.rb_begin(item, part); rb < wk.rb_end(item, part); ++rb)
          for (int ks = 0; ks < (TM / KB) * (rb + 1); ++ks) {
            if (wrapped) mbar_wait(empty + slot, phase);
            if (elect_one()) {
              uint8_t* st = smem + slot * STAGE;
              mbar_arrive_expect_tx(full + slot, STAGE);
              bulk_g2s(st, facS + ((long long)rb * KS_ALL + ks) * A_STAGE, A_STAGE, full + slot);
              bulk_g2s(st + A_STAGE, gB + (long long)ks * B_STAGE, B_STAGE, full + slot);
            }
            __syncwarp();
            if (++slot == NSTAGE) { slot = 0; phase ^= 1u; wrapped = true; }
          }
    }
  } else if (warp == 5) {
    // ---- MMA issuer (warp-converged, one elected lane) ----
    // instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N = 64, M = 128
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    int slot = 0, done = 0;
    unsigned phase = 0;
    for (int item = blockIdx.x; item < wk.n_items(); item += gridDim.x)
      for (int part = 0; part < wk.parts(item); ++part)
        for (int rb = wk.rb_begin(item, part); rb < wk.rb_end(item, part); ++rb, ++done) {
          if (done > 0) { mbar_wait(acc_empty, (unsigned)((done - 1) & 1)); tc_fence_after(); }
          const int n_ks = (TM / KB) * (rb + 1);
          for (int ks = 0; ks < n_ks; ++ks) {
            mbar_wait(full + slot, phase);
            tc_fence_after();
            if (elect_one()) {
              // slice p of L^-1 meets slices 1 .. 8 - p of K* from the collector; class p + q accumulates in its own
              // 64 TMEM columns, and every class sees its first product in the p = 1 group (i8_common.cuh)
              b7i8::gemm::mma_stage(smem + slot * STAGE, tmem, idesc, ks == 0);
              umma_commit(empty + slot);                   // frees the stage once these MMAs have read it
              if (ks == n_ks - 1) umma_commit(acc_full);   // all MMAs of the row block done -> epilogue may read TMEM
            }
            __syncwarp();
            if (++slot == NSTAGE) { slot = 0; phase ^= 1u; }
          }
        }
  } else {
    // ---- epilogue warps 0-3: thread = L^-1 row (TMEM lane), 64 candidate columns ----
    int done = 0;
    for (int item = blockIdx.x; item < wk.n_items(); item += gridDim.x) {
    const int tile = wk.tile(item);
    for (int part = 0; part < wk.parts(item); ++part)
    for (int rb = wk.rb_begin(item, part); rb < wk.rb_end(item, part); ++rb, ++done) {
      mbar_wait(acc_full, (unsigned)(done & 1));
      tc_fence_after();
      double v[TN];
#pragma unroll
      for (int c = 0; c < TN; ++c) v[c] = 0.0;
#pragma unroll
      for (int w = 2; w <= NS + 1; ++w) {
        const double wt = ldexp(1.0, 4 - 8 * w);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t dv[32];
          tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((w - 2) * TN + h * 32), dv);
#pragma unroll
          for (int c = 0; c < 32; ++c) v[h * 32 + c] = fma((double)(int)dv[c], wt, v[h * 32 + c]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);     // TMEM drained: the next row block may start
      const int row = rb * TM + tid;
      const double sc = sigma[row] * tau, b = beta[row];
      double* rr = red + (done & 1) * (4 * TN * 2);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double x[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) { const double vv = v[h * 32 + c] * sc; x[c] = vv * vv; }
        const double t2 = lane_transpose_sum(x, lane);
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = (v[h * 32 + c] * sc) * b;
        const double t1 = lane_transpose_sum(x, lane);
        rr[(warp * TN + h * 32 + lane) * 2 + 0] = t2;
        rr[(warp * TN + h * 32 + lane) * 2 + 1] = t1;
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      if (tid < TN) {
        double* out = partial + (((long long)tile * NB + rb) * TN + tid) * 2;
        out[0] = ((rr[(0 * TN + tid) * 2] + rr[(1 * TN + tid) * 2]) + rr[(2 * TN + tid) * 2]) + rr[(3 * TN + tid) * 2];
        out[1] = ((rr[(0 * TN + tid) * 2 + 1] + rr[(1 * TN + tid) * 2 + 1]) + rr[(2 * TN + tid) * 2 + 1]) + rr[(3 * TN + tid) * 2 + 1];
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

// var = sf2 - sum_rb sum v^2, mean = m + sum_rb sum v beta, row blocks added in ascending order (fixed order:
// deterministic and independent of how the row blocks were split over CTAs)
__global__ void __launch_bounds__(128)
posterior_i8_finish_kernel(const double* __restrict__ partial, int NB, int Np, const double* __restrict__ cand, long long rows, int d,
                           const double* __restrict__ Xt, const double* __restrict__ par, int kernel, double sf2, double mconst,
                           long long cols_pad, double* __restrict__ mean, double* __restrict__ var) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols_pad) return;
  const double* p = partial + ((c / TN) * NB * TN + c % TN) * 2;
  double run2 = 0.0, run1 = 0.0;
  for (int rb = 0; rb < NB; ++rb) {
    run2 += p[(long long)rb * TN * 2];
    run1 += p[(long long)rb * TN * 2 + 1];
  }
  // integers cannot carry a NaN: where the K* row of the fp64 path would be NaN (a NaN coordinate, or an infinite
  // one under Matern: inf * 0), poison the results here.  With finite observations the scaled distance to
  // the first one decides it for the whole row.
  double r2 = 0.0;
  if (c < rows)
    for (int j = 0; j < d; ++j) {
      const double t = (cand[c * d + j] - Xt[(long long)j * Np]) * par[j];
      r2 = fma(t, t, r2);
    }
  const bool nan_row = (r2 != r2) || (kernel == B7_KERNEL_MATERN52 && isinf(r2));
  const double poison = nan_row ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
  const double vv = sf2 - run2;
  var[c] = (vv > 0.0 ? vv : (vv != vv ? vv : 0.0)) + poison;
  mean[c] = mconst + run1 + poison;
}

bool g_attr_i8[16] = {false};

}  // namespace

int b7_i8_slice_factor(b7_ctx* ctx, const double* fac, int Np, int8_t* facS, double* sigma, int s0, int count) {
  slice_factor_kernel<<<dim3(Np / TM, count), SLICE_THREADS, 0, ctx->stream>>>(fac, (long long)Np * Np, Np, facS, (long long)Np * Np * NS, sigma, s0);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

template <int DT>
static int launch_cov_slices(b7_ctx* ctx, cudaStream_t st, int kernel, const double* A, int64_t rows, int64_t rows_pad, int d,
                             const double* Xt, int N, int Np, const double* par, double inv_tau, int8_t* ksS) {
  dim3 grid((unsigned)(rows_pad / TN), Np / 64);
  if (kernel == B7_KERNEL_ARDSE) cov_slices_kernel<DT, B7_KERNEL_ARDSE><<<grid, 128, 0, st>>>(A, rows, d, Xt, N, Np, par, inv_tau, ksS);
  else cov_slices_kernel<DT, B7_KERNEL_MATERN52><<<grid, 128, 0, st>>>(A, rows, d, Xt, N, Np, par, inv_tau, ksS);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

int b7_i8_cov_slices(b7_ctx* ctx, cudaStream_t st, int kernel, const double* A, int64_t rows, int64_t rows_pad, int d, const double* Xt,
                     int N, int Np, const double* par, double tau, int8_t* ksS) {
  if (rows_pad <= 0) return 0;
  const double inv_tau = 1.0 / tau;
  if (d <= 2) return launch_cov_slices<2>(ctx, st, kernel, A, rows, rows_pad, d, Xt, N, Np, par, inv_tau, ksS);
  if (d <= 4) return launch_cov_slices<4>(ctx, st, kernel, A, rows, rows_pad, d, Xt, N, Np, par, inv_tau, ksS);
  if (d <= 6) return launch_cov_slices<6>(ctx, st, kernel, A, rows, rows_pad, d, Xt, N, Np, par, inv_tau, ksS);
  if (d <= 8) return launch_cov_slices<8>(ctx, st, kernel, A, rows, rows_pad, d, Xt, N, Np, par, inv_tau, ksS);
  if (d <= 16) return launch_cov_slices<16>(ctx, st, kernel, A, rows, rows_pad, d, Xt, N, Np, par, inv_tau, ksS);
  if (d <= 24) return launch_cov_slices<24>(ctx, st, kernel, A, rows, rows_pad, d, Xt, N, Np, par, inv_tau, ksS);
  return launch_cov_slices<40>(ctx, st, kernel, A, rows, rows_pad, d, Xt, N, Np, par, inv_tau, ksS);
}

size_t b7_i8_partial_bytes(int Np, int64_t cols_pad) { return (size_t)(cols_pad / TN) * (Np / TM) * TN * 2 * sizeof(double); }

int b7_launch_posterior_i8(b7_ctx* ctx, const int8_t* facS, const double* sigma, const double* beta, int Np, const int8_t* ksS,
                           const double* cand, int64_t rows, int d, const double* Xt, const double* par, int kernel,
                           int64_t cols_pad, double tau, double sf2, double mconst, double* partial, double* mean, double* var) {
  if (!g_attr_i8[ctx->device & 15]) {
    B7_CUDA(cudaFuncSetAttribute(posterior_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    g_attr_i8[ctx->device & 15] = true;
  }
  if (cols_pad <= 0) return 0;
  const int NB = Np / TM;
  static const int chunk_env = getenv("B7_POST_CHUNK") ? atoi(getenv("B7_POST_CHUNK")) : 0;
  // measured at N = 4096 (64 launches, ms): whole tile per CTA 204-206; persistent grid over equal-weight items
  // with chunk 1 / 2 / 4 and groups of 8 / 16 / 32 tiles: 187-194 (differences inside the run-to-run noise)
  const int chunk = chunk_env > 0 ? (chunk_env < NB ? chunk_env : NB) : (NB < 2 ? NB : 2);
  const int cpt = (NB + chunk - 1) / chunk;
  static const int group_env = getenv("B7_POST_GROUP") ? atoi(getenv("B7_POST_GROUP")) : 0;
  const int n_tiles = (int)(cols_pad / TN), group = group_env > 0 ? group_env : 16;
  const int n_items = n_tiles * ((cpt + 1) / 2);
  posterior_i8_kernel<<<n_items < ctx->sm_count ? n_items : ctx->sm_count, I8_THREADS, I8_SMEM, ctx->stream>>>(facS, sigma, beta, Np, NB, ksS, tau,
                                                                                                              chunk, group, n_tiles, partial);
  posterior_i8_finish_kernel<<<(unsigned)((cols_pad + 127) / 128), 128, 0, ctx->stream>>>(partial, NB, Np, cand, rows, d, Xt, par, kernel, sf2,
                                                                                         mconst, cols_pad, mean, var);
  b7_count(ctx, 2);
  B7_CUDA(cudaGetLastError());
  return 0;
}
