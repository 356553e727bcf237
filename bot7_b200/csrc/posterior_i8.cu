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
constexpr int RED_BYTES = 2 * 4 * TN * 8;               // [rb parity][warp][candidate] sum v^2
constexpr int I8_SMEM = NSTAGE * STAGE + RED_BYTES + 1024;   // stages + reduction scratch + barriers
// CTA-pair variant (posterior_i8_pair_kernel): each CTA stages its own 128 rows of L^-1 and HALF of the candidate tile
constexpr int TNH = TN / 2;                // candidates staged per CTA of a pair
constexpr int B_HALF = NS * TNH * KB;      // 14336 B
constexpr int P_STAGE = A_STAGE + B_HALF;  // 71680 B
constexpr int P_NSTAGE = 3;
constexpr int P_SMEM = P_NSTAGE * P_STAGE + RED_BYTES + 1024;   // 220160 B
constexpr int P_THREADS = 320;             // warps 0-7 epilogue (lane quadrant w & 3, column half w >> 2), 8 producer, 9 MMA / relay
static_assert(P_STAGE % 1024 == 0 && P_SMEM <= 227 * 1024, "pair stage geometry");

// ---- slicing -------------------------------------------------------------------------------------------

// Pair-packed slice array of one draw: row blocks 2j and 2j + 1 (the two CTAs of a pair work on them side by
// side) both own 4 (j + 1) stages of A_STAGE bytes -- the even block's last two stages are zeros (its row of
// L^-1 ends one block earlier), as is a phantom last block when the number of blocks is odd.
//   stage offset of row block rb:  4 j (j + 1) + (rb & 1) 4 (j + 1),  j = rb / 2
__host__ __device__ __forceinline__ long long pk_off(int rb) {
  const long long j = rb >> 1;
  return 4 * j * (j + 1) + (rb & 1) * 4 * (j + 1);
}

// L^-1 (tiled fp64, lower) -> per-row power-of-two scale sigma and the slice array
// facS[pk_off(rb) + ks][p][kc][row][16]  (ks = KB-column stage, kc = 16-column chunk inside it)
__global__ void __launch_bounds__(SLICE_THREADS)
slice_factor_kernel(const double* __restrict__ fac, long long fac_stride, int Np, int8_t* __restrict__ facS,
                    long long facS_stride, double* __restrict__ sigma, int s0) {
  __shared__ double s_red[SLICE_THREADS];
  const int rb = blockIdx.x, s = s0 + blockIdx.y, row = threadIdx.x & 127, q = threadIdx.x >> 7;
  const int KTA = Np / 16, NB = Np / TM;
  const int n_kc = rb < NB ? (rb + 1) * (TM / 16) : 0, all_kc = ((rb >> 1) + 1) * 4 * KC;
  int8_t* dst = facS + (long long)s * facS_stride + pk_off(rb) * A_STAGE;
  double inv = 0.0;
  const double* src = fac + (long long)s * fac_stride + b7g::tile_off(KTA, rb < NB ? rb : 0, 0);
  if (rb < NB) {                                    // block-uniform
    double mx = 0.0;
    bool bad = false;
    for (int kc = q; kc < n_kc; kc += 4) chunk_max(src + b7g::elem_off(row, kc * 16), mx, bad);
    // a failed factorisation (NaN / inf in the row) must poison the results like it does on the fp64 path
    const double sg = row_scale(mx, bad, s_red, row, q);
    inv = sg != sg ? 0.0 : 1.0 / sg;
    if (q == 0) sigma[(long long)s * Np + rb * TM + row] = sg;
  }
  for (int kc = q; kc < all_kc; kc += 4) {
    uint32_t pk[4][NS];
    if (kc < n_kc) {
      chunk_digits(src + b7g::elem_off(row, kc * 16), inv, pk);
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int p = 0; p < NS; ++p) pk[g][p] = 0u;
    }
    const int ks = kc / KC, kcc = kc % KC;
#pragma unroll
    for (int p = 0; p < NS; ++p)
      *reinterpret_cast<uint4*>(dst + (long long)ks * A_STAGE + p * (KC * TM * 16) + kcc * (TM * 16) + row * 16) =
          make_uint4(pk[0][p], pk[1][p], pk[2][p], pk[3][p]);
  }
}

// K(X*, X) evaluated and sliced in one pass, one block per (candidate tile, 64 columns); 128 threads, two 16-column
// chunks each (21.6 -> 16.9 ms per step against 256 threads x one chunk, together with folding sf2 / tau * 2^54 into
// one factor).  Slice layout: PAIR = false  ksS[ct][ks][q][kc][cand 64][16]   (one CTA stages the whole tile),
//                             PAIR = true   ksS[ct][ks][half][q][kc][cand 32][16]   (each CTA of a pair stages a half).
// The same pass leaves the fp64 dot product of its 64 columns with alpha, meanP[ct][kb64][cand] = sum_k k*(c, k)
// alpha_k * 2^54 / tau: the posterior mean m + k*^T alpha is then formed from unsliced fp64 values, the way the
// oracle forms it, and only the variance goes through the int8 products.
template <int DT, int KERNEL, bool PAIR>
__global__ void __launch_bounds__(128)
cov_slices_kernel(const double* __restrict__ A, long long rows, int d, const double* __restrict__ Xt, int N, int Np,
                  const double* __restrict__ par, double inv_tau, const double* __restrict__ alpha, int8_t* __restrict__ ksS,
                  double* __restrict__ meanP) {
  __shared__ __align__(16) double s_x[DT][64];
  __shared__ double s_w[DT];
  __shared__ double s_tab[64];
  __shared__ __align__(16) double s_al[64];
  __shared__ double s_dot[64];
  if (threadIdx.x < 64) s_tab[threadIdx.x] = c_exp_tab[threadIdx.x];
  const int ct = blockIdx.x, kb64 = blockIdx.y, KS_ALL = Np / KB;
  const int c = threadIdx.x & 63;                                 // warp = 32 consecutive candidates, one k chunk
  for (int e = threadIdx.x; e < DT * 64; e += 128) {
    const int i = e / 64, k = kb64 * 64 + e % 64;
    s_x[i][e % 64] = (i < d && k < N) ? Xt[(long long)i * Np + k] : 0.0;
  }
  if (threadIdx.x >= 64) s_al[c] = alpha[kb64 * 64 + c];
  if (threadIdx.x < DT) s_w[threadIdx.x] = threadIdx.x < d ? par[threadIdx.x] : 0.0;
  __syncthreads();
  const long long row = (long long)ct * TN + c;
  double a[DT];
#pragma unroll
  for (int i = 0; i < DT; ++i) a[i] = (row < rows && i < d) ? A[row * d + i] : 0.0;
  // sf2 / tau * 2^54: tau is a power of two, so scaling sf2 first rounds exactly like scaling the product afterwards
  const double sf2s = par[B7_MAX_DIMS] * inv_tau * 18014398509481984.0;
  double dot = 0.0;
  double w[DT];
#pragma unroll
  for (int j = 0; j < DT; ++j) w[j] = s_w[j];
  ExpNegK K;
  K.init();
  const bool live = row < rows;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
  const int kc = (threadIdx.x >> 6) + 2 * half;
  uint32_t pk[4][NS];
  unsigned long long z[4];
  // two observations per step (one LDS.128 per dimension); the bounds are selects, not branches (k is warp-uniform, the
  // padded columns of s_x are zeros, a dead row computes on zeros): the kernel is bound by its instruction count
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const int kk = kc * 16 + i, k = kb64 * 64 + kk;
    double r2a = 0.0, r2b = 0.0;
#pragma unroll
    for (int j = 0; j < DT; ++j) {
      const double2 x2 = *reinterpret_cast<const double2*>(&s_x[j][kk]);
      const double ta = (a[j] - x2.x) * w[j], tb = (a[j] - x2.y) * w[j];
      r2a = fma(ta, ta, r2a);
      r2b = fma(tb, tb, r2b);
    }
    double va, vb;
    if (KERNEL == B7_KERNEL_ARDSE) {
      va = sf2s * exp_neg_k(-0.5 * r2a, s_tab, K);
      vb = sf2s * exp_neg_k(-0.5 * r2b, s_tab, K);
    } else {
      const double ra = sqrt(r2a), s5a = 2.23606797749978969641 * ra, rb = sqrt(r2b), s5b = 2.23606797749978969641 * rb;
      va = sf2s * ((1.0 + s5a + (5.0 / 3.0) * r2a) * exp_neg_k(-s5a, s_tab, K));
      vb = sf2s * ((1.0 + s5b + (5.0 / 3.0) * r2b) * exp_neg_k(-s5b, s_tab, K));
    }
    va = (live && k < N) ? va : 0.0;
    vb = (live && k + 1 < N) ? vb : 0.0;
    const double2 al = *reinterpret_cast<const double2*>(&s_al[kk]);
    dot = fma(va, al.x, dot);
    dot = fma(vb, al.y, dot);
    z[i & 3] = digit_bytes_scaled(va);
    z[(i & 3) + 1] = digit_bytes_scaled(vb);
    if ((i & 3) == 2) pack4(z, pk[i >> 2]);
  }
  const int gkc = kb64 * 4 + kc, ks = gkc / KC, kcc = gkc % KC;
  int8_t* dst = PAIR ? ksS + (((long long)ct * KS_ALL + ks) * 2 + (c >> 5)) * B_HALF + kcc * (TNH * 16) + (c & 31) * 16
                     : ksS + ((long long)ct * KS_ALL + ks) * B_STAGE + kcc * (TN * 16) + c * 16;
  constexpr int P_STRIDE = PAIR ? KC * TNH * 16 : KC * TN * 16;
#pragma unroll
  for (int p = 0; p < NS; ++p) *reinterpret_cast<uint4*>(dst + p * P_STRIDE) = make_uint4(pk[0][p], pk[1][p], pk[2][p], pk[3][p]);
  }
  // chunks 0, 2 (threads 0-63) + chunks 1, 3 (threads 64-127), fixed order
  if (threadIdx.x >= 64) s_dot[c] = dot;
  __syncthreads();
  if (threadIdx.x < 64) meanP[((long long)ct * (Np / 64) + kb64) * 64 + c] = dot + s_dot[c];
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

// ---- the kernels ---------------------------------------------------------------------------------------

// epilogue of one row block (warps 0-3 of either kernel, thread = L^-1 row = TMEM lane): rebuild v in fp64 from the
// 7 class sums, then sum (v sigma_row tau)^2 over the 128 rows; one partial per candidate lands in out[0..63]
// (out == nullptr: a phantom row block, nothing to store).  `arrive_drained` runs once the accumulators are in registers.
template <typename F>
__device__ __forceinline__ void epilogue_row_block(uint32_t tmem, int tid, int warp, int lane, double sc, double* rr, double* out,
                                                   F arrive_drained) {
  double v[TN];
  b7i8::gemm::drain_classes(tmem, warp, v);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) arrive_drained();               // TMEM drained: the next row block may start
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) { const double vv = v[h * 32 + c] * sc; x[c] = vv * vv; }
    rr[warp * TN + h * 32 + lane] = lane_transpose_sum(x, lane);
  }
  asm volatile("bar.sync 1, 128;\n" ::: "memory");
  if (tid < TN && out) out[tid] = ((rr[0 * TN + tid] + rr[1 * TN + tid]) + rr[2 * TN + tid]) + rr[3 * TN + tid];
}

// The pair kernel's epilogue: 8 warps, warp w drains TMEM lanes 32 (w & 3) .. and the 32 candidate columns of half
// w >> 2; the 7 class loads go out in two batches (3 + 4) with one wait per batch (the one-load-one-wait drain of
// 4 warps kept the accumulators busy for 2.9 us per row block, 6.6 % of the kernel: ncu pair_r02).  Same fma chain
// per entry and same summation tree as epilogue_row_block: bit-identical results.
template <typename F>
__device__ __forceinline__ void epilogue_row_block8(uint32_t tmem, int tid, int warp, int lane, double sc, double* rr, double* out,
                                                    F arrive_drained) {
  const int q = warp & 3, hc = warp >> 2;
  const uint32_t base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hc * 32);
  double v[32];
  uint32_t a[4][32];
  // classes 0-2, then classes 3-6: the accumulators are released as soon as the second batch sits in registers, before its
  // int -> fp64 conversion (4/7 of the FP64 work of the drain leaves the MMA warp's critical path)
#pragma unroll
  for (int i = 0; i < 3; ++i) tmem_ld32_async(base + (uint32_t)(i * TN), a[i]);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 32; ++c) v[c] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    tmem_pin(a[i]);
    const double wt = ldexp(1.0, 4 - 8 * (i + 2));
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = fma((double)(int)a[i][c], wt, v[c]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) tmem_ld32_async(base + (uint32_t)((3 + i) * TN), a[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i) tmem_pin(a[i]);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) arrive_drained();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double wt = ldexp(1.0, 4 - 8 * (i + 5));
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = fma((double)(int)a[i][c], wt, v[c]);
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) { const double vv = v[c] * sc; v[c] = vv * vv; }
  rr[q * TN + hc * 32 + lane] = lane_transpose_sum(v, lane);
  asm volatile("bar.sync 1, 256;\n" ::: "memory");
  if (tid < TN && out) out[tid] = ((rr[0 * TN + tid] + rr[1 * TN + tid]) + rr[2 * TN + tid]) + rr[3 * TN + tid];
}

__global__ void __launch_bounds__(I8_THREADS, 1)
posterior_i8_kernel(const int8_t* __restrict__ facS, const double* __restrict__ sigma, int Np, int NB,
                    const int8_t* __restrict__ ksS, double tau, int chunk, int group, int n_tiles, double* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem[];
  double* red = reinterpret_cast<double*>(smem + NSTAGE * STAGE);             // [2 parity][4 warps][64 cols]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE + RED_BYTES);
  uint64_t *full = bars, *empty = bars + NSTAGE, *acc_full = bars + 2 * NSTAGE, *acc_empty = bars + 2 * NSTAGE + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS_ALL = Np / KB;
  // Work items of equal weight: a 64-candidate tile x a pair of row-block chunks (chunk l and chunk cpt-1-l: the
  // triangular L^-1 makes their stage counts add up to the same number for every l).  Each item leaves one partial
  // sum v^2 per candidate and row block; posterior_i8_finish_kernel adds them in row-block order.
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
        for (int rb = wk.rb_begin(item, part); rb < wk.rb_end(item, part); ++rb) {
          const int8_t* gA = facS + pk_off(rb) * A_STAGE;
          for (int ks = 0; ks < (TM / KB) * (rb + 1); ++ks) {
            if (wrapped) mbar_wait(empty + slot, phase);
            if (elect_one()) {
              uint8_t* st = smem + slot * STAGE;
              mbar_arrive_expect_tx(full + slot, STAGE);
              bulk_g2s(st, gA + (long long)ks * A_STAGE, A_STAGE, full + slot);
              bulk_g2s(st + A_STAGE, gB + (long long)ks * B_STAGE, B_STAGE, full + slot);
            }
            __syncwarp();
            if (++slot == NSTAGE) { slot = 0; phase ^= 1u; wrapped = true; }
          }
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
    // ---- epilogue warps 0-3 ----
    int done = 0;
    for (int item = blockIdx.x; item < wk.n_items(); item += gridDim.x) {
      const int tile = wk.tile(item);
      for (int part = 0; part < wk.parts(item); ++part)
        for (int rb = wk.rb_begin(item, part); rb < wk.rb_end(item, part); ++rb, ++done) {
          mbar_wait(acc_full, (unsigned)(done & 1));
          tc_fence_after();
          epilogue_row_block(tmem, tid, warp, lane, sigma[rb * TM + tid] * tau, red + (done & 1) * (4 * TN),
                             partial + ((long long)tile * NB + rb) * TN, [&] { mbar_arrive(acc_empty); });
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

// The same pass on CTA pairs (tcgen05 cta_group::2): the two CTAs of a cluster take row blocks 2j and 2j + 1 of L^-1
// against the SAME 64-candidate tile.  One instruction of the leader drives both tensor cores (M = 256); each CTA
// stages its own 128 rows of L^-1 but only 32 of the 64 candidates, and the hardware reads the two halves of the
// column operand from both shared memories: 71.7 KB instead of 86 KB per CTA-stage from L2, half the shared-memory
// operand reads for K* per SM, and room for a third stage.  Protocol per stage: both producers fill their own slot
// (own "full" barrier); warp 5 of the peer relays its "full" to the leader (remote arrive on peer_full); the leader's
// MMA warp waits for both, issues, and commits with a multicast arrive that frees the slot in both CTAs; "accumulators
// ready" is a multicast commit as well; the 8 epilogue warps of the pair release the accumulators on the leader's
// acc_empty.  Work items: a tile x pair-chunks (l, P-1-l), P = ceil(NB / 2), 4 (P + 1) stages each.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
posterior_i8_pair_kernel(const int8_t* __restrict__ facS, const double* __restrict__ sigma, int Np, int NB,
                         const int8_t* __restrict__ ksS, double tau, int group, int n_tiles, double* __restrict__ partial, int dbg) {
  // dbg (B7_POST_DBG, measurement only, results are then wrong): 1 = epilogue skips the drain and the fp64 work,
  // 2 = producers always fetch the first stage of the first tile (every copy hits the same hot L2 lines),
  // 4 = producers fetch 2 KB per stage (no operand traffic to speak of)
  extern __shared__ __align__(1024) uint8_t smem[];
  double* red = reinterpret_cast<double*>(smem + P_NSTAGE * P_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_NSTAGE * P_STAGE + RED_BYTES);
  uint64_t *full = bars, *empty = bars + P_NSTAGE, *peer_full = bars + 2 * P_NSTAGE, *acc_full = bars + 3 * P_NSTAGE,
           *acc_empty = bars + 3 * P_NSTAGE + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * P_NSTAGE + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS_ALL = Np / KB;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const Walk wk{(NB + 1) / 2, 1, group, n_tiles};     // "row block" of the walk = pair-chunk j = row blocks 2j, 2j + 1
  if (tid == 0) {
    for (int s = 0; s < P_NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); mbar_init(peer_full + s, 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 16);                         // 8 epilogue warps in each CTA of the pair
    mbar_fence_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // barriers and TMEM of both CTAs exist before anything crosses
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ---- producer (both CTAs): own row block + own half of the candidate tile ----
    const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
    int slot = 0;
    unsigned phase = 1;
    bool wrapped = false;
    for (int item = pair_id; item < wk.n_items(); item += n_pairs) {
      const int tile = wk.tile(item);
      const int8_t* gB = ksS + ((long long)tile * KS_ALL * 2 + rank) * B_HALF;
      for (int part = 0; part < wk.parts(item); ++part)
        for (int j = wk.rb_begin(item, part); j < wk.rb_end(item, part); ++j) {
          const int8_t* gA = facS + pk_off(2 * j + (int)rank) * A_STAGE;
          for (int ks = 0; ks < 4 * (j + 1); ++ks) {
            if (wrapped) mbar_wait_cluster(empty + slot, phase);
            if (elect_one()) {
              uint8_t* st = smem + slot * P_STAGE;
              if (dbg & 4) {                         // 2 KB instead of 71.7 KB per stage: the MMAs run on stale operands
                mbar_arrive_expect_tx(full + slot, 2048);
                bulk_g2s(st, facS, 1024, full + slot);
                bulk_g2s(st + A_STAGE, ksS, 1024, full + slot);
                __syncwarp();
                if (++slot == P_NSTAGE) { slot = 0; phase ^= 1u; wrapped = true; }
                continue;
              }
              mbar_arrive_expect_tx(full + slot, P_STAGE);
              if (dbg & 2) {
                bulk_g2s(st, facS, A_STAGE, full + slot);
                bulk_g2s(st + A_STAGE, ksS, B_HALF, full + slot);
              } else {
                // measurement knob: 1 = L^-1 evict_last, 2 = + K* evict_first, 3 = K* evict_first only,
                //                   4 = L^-1 evict_first + K* evict_last, 5 = L^-1 evict_first only, 6 = K* evict_last only
                const int hint = dbg >> 4;
                if (hint == 1 || hint == 2) bulk_g2s_hint(st, gA + (long long)ks * A_STAGE, A_STAGE, full + slot, pol_last);
                else if (hint == 4 || hint == 5) bulk_g2s_hint(st, gA + (long long)ks * A_STAGE, A_STAGE, full + slot, pol_first);
                else bulk_g2s(st, gA + (long long)ks * A_STAGE, A_STAGE, full + slot);
                if (hint == 2 || hint == 3) bulk_g2s_hint(st + A_STAGE, gB + (long long)ks * (2 * B_HALF), B_HALF, full + slot, pol_first);
                else if (hint == 4 || hint == 6) bulk_g2s_hint(st + A_STAGE, gB + (long long)ks * (2 * B_HALF), B_HALF, full + slot, pol_last);
                else bulk_g2s(st + A_STAGE, gB + (long long)ks * (2 * B_HALF), B_HALF, full + slot);
              }
            }
            __syncwarp();
            if (++slot == P_NSTAGE) { slot = 0; phase ^= 1u; wrapped = true; }
          }
        }
    }
  } else if (warp == 9) {
    int slot = 0, done = 0;
    unsigned phase = 0;
    if (rank == 0) {
      // ---- MMA issuer (leader): D = S32, A = B = signed 8 bit, K-major, N = 64, M = 256 (the pair) ----
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      for (int item = pair_id; item < wk.n_items(); item += n_pairs)
        for (int part = 0; part < wk.parts(item); ++part)
          for (int j = wk.rb_begin(item, part); j < wk.rb_end(item, part); ++j, ++done) {
            if (done > 0) { mbar_wait_cluster(acc_empty, (unsigned)((done - 1) & 1)); tc_fence_after(); }
            const int n_ks = 4 * (j + 1);
            for (int ks = 0; ks < n_ks; ++ks) {
              mbar_wait_cluster(full + slot, phase);
              mbar_wait_cluster(peer_full + slot, phase);
              tc_fence_after();
              if (elect_one()) {
                const uint8_t* st = smem + slot * P_STAGE;
                const uint64_t da0 = umma_desc(st, TM * 16, 128), db0 = umma_desc(st + A_STAGE, TNH * 16, 128);
#pragma unroll
                for (int k2 = 0; k2 < KB / 32; ++k2) {
#pragma unroll
                  for (int p = 1; p <= NS; ++p) {
                    const uint64_t da = da0 + (uint64_t)(((p - 1) * (KC * TM * 16) + k2 * (2 * TM * 16)) >> 4);
                    const uint32_t acc = (ks == 0 && k2 == 0 && p == 1) ? 0u : 1u;
                    const int nq = NS + 1 - p;
#pragma unroll
                    for (int q = 1; q <= nq; ++q) {
                      const uint64_t db = db0 + (uint64_t)(((q - 1) * (KC * TNH * 16) + k2 * (2 * TNH * 16)) >> 4);
                      const uint32_t dcol = tmem + (uint32_t)((p + q - 2) * TN);
                      if (nq == 1) umma_i8_pair<0>(dcol, da, db, idesc, acc);
                      else if (q == 1) umma_i8_pair<1>(dcol, da, db, idesc, acc);
                      else if (q == nq) umma_i8_pair<3>(dcol, da, db, idesc, acc);
                      else umma_i8_pair<2>(dcol, da, db, idesc, acc);
                    }
                  }
                }
                umma_commit_pair(empty + slot);                   // frees the slot in both CTAs
                if (ks == n_ks - 1) umma_commit_pair(acc_full);   // both CTAs' epilogues may read their TMEM
              }
              __syncwarp();
              if (++slot == P_NSTAGE) { slot = 0; phase ^= 1u; }
            }
          }
    } else {
      // ---- relay (peer): "my slot is full" -> the leader's peer_full ----
      for (int item = pair_id; item < wk.n_items(); item += n_pairs)
        for (int part = 0; part < wk.parts(item); ++part)
          for (int j = wk.rb_begin(item, part); j < wk.rb_end(item, part); ++j)
            for (int ks = 0; ks < 4 * (j + 1); ++ks) {
              mbar_wait_cluster(full + slot, phase);
              if (elect_one()) mbar_arrive_cluster(peer_full + slot, 0);
              __syncwarp();
              if (++slot == P_NSTAGE) { slot = 0; phase ^= 1u; }
            }
    }
  } else {
    // ---- epilogue warps 0-7 of both CTAs: CTA `rank` holds row block 2j + rank in its TMEM ----
    int done = 0;
    for (int item = pair_id; item < wk.n_items(); item += n_pairs) {
      const int tile = wk.tile(item);
      for (int part = 0; part < wk.parts(item); ++part)
        for (int j = wk.rb_begin(item, part); j < wk.rb_end(item, part); ++j, ++done) {
          const int rb = 2 * j + (int)rank;
          mbar_wait_cluster(acc_full, (unsigned)(done & 1));
          tc_fence_after();
          if (dbg & 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty, 0);
            continue;
          }
          const double sc = rb < NB ? sigma[rb * TM + (tid & 127)] * tau : 0.0;
          epilogue_row_block8(tmem, tid, warp, lane, sc, red + (done & 1) * (4 * TN),
                             rb < NB ? partial + ((long long)tile * NB + rb) * TN : nullptr, [&] { mbar_arrive_cluster(acc_empty, 0); });
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // nobody leaves while the peer may still read its shared memory
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

// var = sf2 - sum_rb sum v^2 (row blocks added in ascending order), mean = m + tau 2^-54 sum_kb meanP (column blocks
// in ascending order): fixed orders, deterministic and independent of how the work was split over CTAs
__global__ void __launch_bounds__(128)
posterior_i8_finish_kernel(const double* __restrict__ partial, const double* __restrict__ meanP, int NB, int Np,
                           const double* __restrict__ cand, long long rows, int d, const double* __restrict__ Xt,
                           const double* __restrict__ par, int kernel, double sf2, double mconst, double mean_scale,
                           long long cols_pad, double* __restrict__ mean, double* __restrict__ var) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols_pad) return;
  const double* p = partial + (c / TN) * NB * TN + c % TN;
  double run2 = 0.0, run1 = 0.0;
  for (int rb = 0; rb < NB; ++rb) run2 += p[(long long)rb * TN];
  const double* q = meanP + (c / TN) * (Np / 64) * 64 + c % TN;
  for (int kb = 0; kb < Np / 64; ++kb) run1 += q[(long long)kb * 64];
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
  mean[c] = (nan_row ? 0.0 : fma(run1, mean_scale, mconst)) + poison;
}

bool g_attr_i8[16] = {false};

}  // namespace

size_t b7_i8_facs_stride(int Np) {
  const long long P = (Np / TM + 1) / 2;
  return (size_t)(4 * P * (P + 1)) * A_STAGE;
}

int b7_i8_slice_factor(b7_ctx* ctx, const double* fac, int Np, int8_t* facS, double* sigma, int s0, int count) {
  const int NBp = (Np / TM + 1) / 2 * 2;             // a phantom last block (zeros) when the number of blocks is odd
  slice_factor_kernel<<<dim3(NBp, count), SLICE_THREADS, 0, ctx->stream>>>(fac, (long long)Np * Np, Np, facS, (long long)b7_i8_facs_stride(Np), sigma, s0);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

template <int DT>
static int launch_cov_slices(b7_ctx* ctx, cudaStream_t st, int kernel, const double* A, int64_t rows, int64_t rows_pad, int d,
                             const double* Xt, int N, int Np, const double* par, double inv_tau, const double* alpha, int8_t* ksS,
                             double* meanP) {
  dim3 grid((unsigned)(rows_pad / TN), Np / 64);
  const bool se = kernel == B7_KERNEL_ARDSE;
  if (ctx->post_pair) {
    if (se) cov_slices_kernel<DT, B7_KERNEL_ARDSE, true><<<grid, 128, 0, st>>>(A, rows, d, Xt, N, Np, par, inv_tau, alpha, ksS, meanP);
    else cov_slices_kernel<DT, B7_KERNEL_MATERN52, true><<<grid, 128, 0, st>>>(A, rows, d, Xt, N, Np, par, inv_tau, alpha, ksS, meanP);
  } else {
    if (se) cov_slices_kernel<DT, B7_KERNEL_ARDSE, false><<<grid, 128, 0, st>>>(A, rows, d, Xt, N, Np, par, inv_tau, alpha, ksS, meanP);
    else cov_slices_kernel<DT, B7_KERNEL_MATERN52, false><<<grid, 128, 0, st>>>(A, rows, d, Xt, N, Np, par, inv_tau, alpha, ksS, meanP);
  }
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

int b7_i8_cov_slices(b7_ctx* ctx, cudaStream_t st, int kernel, const double* A, int64_t rows, int64_t rows_pad, int d, const double* Xt,
                     int N, int Np, const double* par, double tau, const double* alpha, int8_t* ksS, double* meanP) {
  if (rows_pad <= 0) return 0;
  const double inv_tau = 1.0 / tau;
#define B7_COV_SLICES(DT) return launch_cov_slices<DT>(ctx, st, kernel, A, rows, rows_pad, d, Xt, N, Np, par, inv_tau, alpha, ksS, meanP)
  if (d <= 2) B7_COV_SLICES(2);
  if (d <= 4) B7_COV_SLICES(4);
  if (d <= 6) B7_COV_SLICES(6);
  if (d <= 8) B7_COV_SLICES(8);
  if (d <= 12) B7_COV_SLICES(12);
  if (d <= 16) B7_COV_SLICES(16);
  if (d <= 20) B7_COV_SLICES(20);
  if (d <= 24) B7_COV_SLICES(24);
  if (d <= 32) B7_COV_SLICES(32);
  B7_COV_SLICES(40);
#undef B7_COV_SLICES
}

// scratch of one posterior launch: [tile][row block][64] sum v^2 partials, then [tile][64-column block][64] mean partials
size_t b7_i8_partial_bytes(int Np, int64_t cols_pad) {
  return (size_t)(cols_pad / TN) * ((size_t)(Np / TM) * TN + (size_t)(Np / 64) * 64) * sizeof(double);
}
double* b7_i8_mean_partials(double* partial, int Np, int64_t cols_pad) { return partial + (size_t)(cols_pad / TN) * (Np / TM) * TN; }

int b7_launch_posterior_i8(b7_ctx* ctx, const int8_t* facS, const double* sigma, int Np, const int8_t* ksS, const double* cand,
                           int64_t rows, int d, const double* Xt, const double* par, int kernel, int64_t cols_pad, double tau,
                           double sf2, double mconst, double* partial, double* mean, double* var) {
  if (!g_attr_i8[ctx->device & 15]) {
    B7_CUDA(cudaFuncSetAttribute(posterior_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    B7_CUDA(cudaFuncSetAttribute(posterior_i8_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    g_attr_i8[ctx->device & 15] = true;
  }
  if (cols_pad <= 0) return 0;
  const int NB = Np / TM;
  static const int group_env = getenv("B7_POST_GROUP") ? atoi(getenv("B7_POST_GROUP")) : 0;
  const int n_tiles = (int)(cols_pad / TN), group = group_env > 0 ? group_env : 16;
  if (ctx->post_pair) {
    const int P = (NB + 1) / 2, n_items = n_tiles * ((P + 1) / 2), n_pairs = ctx->sm_count / 2;
    static const int dbg_env = getenv("B7_POST_DBG") ? atoi(getenv("B7_POST_DBG")) : 0;
    posterior_i8_pair_kernel<<<2 * (n_items < n_pairs ? n_items : n_pairs), P_THREADS, P_SMEM, ctx->stream>>>(facS, sigma, Np, NB, ksS, tau, group,
                                                                                                              n_tiles, partial, dbg_env);
  } else {
    static const int chunk_env = getenv("B7_POST_CHUNK") ? atoi(getenv("B7_POST_CHUNK")) : 0;
    // measured at N = 4096 (64 launches, ms): whole tile per CTA 204-206; persistent grid over equal-weight items
    // with chunk 1 / 2 / 4 and groups of 8 / 16 / 32 tiles: 187-194 (differences inside the run-to-run noise)
    const int chunk = chunk_env > 0 ? (chunk_env < NB ? chunk_env : NB) : (NB < 2 ? NB : 2);
    const int cpt = (NB + chunk - 1) / chunk;
    const int n_items = n_tiles * ((cpt + 1) / 2);
    posterior_i8_kernel<<<n_items < ctx->sm_count ? n_items : ctx->sm_count, I8_THREADS, I8_SMEM, ctx->stream>>>(facS, sigma, Np, NB, ksS, tau, chunk,
                                                                                                                group, n_tiles, partial);
  }
  // k* = val tau 2^-54 (exact power-of-two factor taken out of the dot product)
  posterior_i8_finish_kernel<<<(unsigned)((cols_pad + 127) / 128), 128, 0, ctx->stream>>>(partial, b7_i8_mean_partials(partial, Np, cols_pad), NB, Np,
                                                                                         cand, rows, d, Xt, par, kernel, sf2, mconst,
                                                                                         tau * 5.5511151231257827e-17, cols_pad, mean, var);
  b7_count(ctx, 2);
  B7_CUDA(cudaGetLastError());
  return 0;
}
