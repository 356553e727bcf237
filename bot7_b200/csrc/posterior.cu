// Posterior over a candidate panel (the FP64-dense part of kernel #2/#3):
//   V = L^-1 K*^T,  var_c = max(sf2 - sum_i V_ic^2, 0),  mean_c = m + sum_i V_ic beta_i,  beta = L^-1 (y - m)
// (mean = m + k*^T alpha with alpha = L^-T beta, written so that it falls out of the same product).
//
// Replaces the triangular solve + column-sum-of-squares inside gpTorch7's gp_regressor:predict
// (external; called at reference scores/expected_improvement.lua:63) -- N^2 flop per candidate per
// draw, ~99 % of the flops of the whole path.
//
// L^-1 is explicit (potrf.cu), so V is a triangular GEMM with no sequential dependency.  One CTA
// owns 128 candidates and walks down all row blocks of L^-1: for row block rb it accumulates the
// 128x128 tile V[rb] over k < (rb+1)*128 on DMMA tiles, folds the tile into per-candidate
// sum(v^2) and sum(v*beta) held in registers, and moves on.  V is never written.
//
// Data movement: both operands are kept in HBM in the fragment order of gemm_tile.cuh ("tiled
// layout": potrf.cu keeps the factor in it throughout, the K* builder writes it directly), so one k-step of an
// operand is 16 contiguous KB.  One thread issues one cp.async.bulk (TMA unit) per operand per
// stage and arms an mbarrier with the byte count; the 8 DMMA warps wait on that "full" barrier,
// compute, and release the slot through an "empty" barrier.  There is no CTA-wide
// barrier and no per-thread copy instruction in the main loop: measured on B200 the same loop fed by
// per-thread cp.async (LDGSTS) lost 22 % of the DMMA issue slots to LSU contention
// (tools/dmma_loop.cu: 36.97 TFLOP/s from resident shared memory, 28.8 with the LDGSTS stream).
// All CTAs execute the same schedule, so the L^-1 tiles they share hit in L2, and a grid of 148 CTAs
// is one full wave with no tail.  Summation order is fixed: run-to-run deterministic.
#include "b7_internal.h"
#include "gemm_tile.cuh"

using namespace b7g;

namespace {

static_assert(BK == TILE_K, "posterior_kernel assumes 16-wide k-tiles");
constexpr int P_STAGES = 6;                                    // 6 x 32 KB
constexpr int P_AHEAD = P_STAGES - 2;                          // stages in flight ahead of the consumer
constexpr int P_THREADS = THREADS;                             // 8 DMMA warps; lane 0 of warp 0 also feeds the ring
constexpr int P_SMEM = P_STAGES * 2 * TILE_DOUBLES * 8 + 1024; // + barriers

__global__ void __launch_bounds__(P_THREADS, 1)
posterior_kernel(const double* __restrict__ LinvT, const double* __restrict__ beta, int Np, int NB,
                 const double* __restrict__ ksT, double sf2, double mconst, double* __restrict__ mean,
                 double* __restrict__ var) {
  extern __shared__ __align__(128) double smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + P_STAGES * 2 * TILE_DOUBLES);
  uint64_t* empty = full + P_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KPB = B7_NB / TILE_K;            // k-tiles per 128-block (8)
  const int KT_ALL = Np / TILE_K;            // k-tiles per matrix row block
  const long long total = (long long)KPB * NB * (NB + 1) / 2;
  if (tid == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
    mbar_fence_init();
  }
  __syncthreads();

  // producer state (only meaningful in thread 0): next tile of the (rb, kt) sequence to fetch.
  // A 9th warp cannot be afforded (register allocation is per 4 warps: 288 threads x 208 registers
  // is rejected), so thread 0 issues the two bulk copies of stage it + P_AHEAD at the top of its
  // iteration it.  The slot it refills was used by stage it - 2, i.e. the wait on its "empty"
  // barrier only blocks if some warp is more than two stages behind.
  const double* gB = ksT + (long long)blockIdx.x * KT_ALL * TILE_DOUBLES;
  int p_rb = 0, p_kt = 0;
  long long p_it = 0;
  auto produce = [&]() {
    const int slot = (int)(p_it % P_STAGES);
    if (p_it >= P_STAGES) mbar_wait(empty + slot, (unsigned)(((p_it / P_STAGES) - 1) & 1));
    double* st = smem + slot * 2 * TILE_DOUBLES;
    mbar_arrive_expect_tx(full + slot, 2 * TILE_DOUBLES * 8);
    bulk_g2s(st, LinvT + ((long long)p_rb * KT_ALL + p_kt) * TILE_DOUBLES, TILE_DOUBLES * 8, full + slot);
    bulk_g2s(st + TILE_DOUBLES, gB + (long long)p_kt * TILE_DOUBLES, TILE_DOUBLES * 8, full + slot);
    if (++p_kt == (p_rb + 1) * KPB) { p_kt = 0; ++p_rb; }
    ++p_it;
  };
  if (tid == 0)
    for (int s = 0; s < P_AHEAD && p_it < total; ++s) produce();

  // ---- consumers: 8 DMMA warps ----
  const int wm = warp >> 2, wn = warp & 3;
  Acc acc; acc.zero();
  double sum2[4][2], sum1[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) sum2[j][0] = sum2[j][1] = sum1[j][0] = sum1[j][1] = 0.0;
  int c_rb = 0, c_kt = 0;
  for (long long it = 0; it < total; ++it) {
    if (tid == 0 && p_it < total) produce();
    const int slot = (int)(it % P_STAGES);
    mbar_wait(full + slot, (unsigned)((it / P_STAGES) & 1));
    const double* st = smem + slot * 2 * TILE_DOUBLES;
    if (c_kt >= c_rb * KPB) compute_stage_tri(st, st + TILE_DOUBLES, wm, wn, lane, acc, (c_kt - c_rb * KPB) * TILE_K);
    else compute_stage(st, st + TILE_DOUBLES, wm, wn, lane, acc);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + slot);
    if (++c_kt == (c_rb + 1) * KPB) {
      // fold the finished V tile of row block c_rb
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double b = __ldg(beta + c_rb * B7_NB + frag_row(wm, i, lane));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double v0 = acc.c[i][j][0], v1 = acc.c[i][j][1];
          sum2[j][0] = fma(v0, v0, sum2[j][0]);
          sum2[j][1] = fma(v1, v1, sum2[j][1]);
          sum1[j][0] = fma(v0, b, sum1[j][0]);
          sum1[j][1] = fma(v1, b, sum1[j][1]);
          acc.c[i][j][0] = 0.0;
          acc.c[i][j][1] = 0.0;
        }
      }
      c_kt = 0;
      ++c_rb;
    }
  }
  // reduce over the 8 row lanes of the warp (lane >> 2), then over the two row warps (wm)
  double* red = smem;   // reuse stage 0 ([2][128][2]): every stage has been consumed once all 8 warps are here
  asm volatile("bar.sync 1, 256;\n" ::: "memory");
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double a = sum2[j][h], b = sum1[j][h];
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
      }
      if ((lane >> 2) == 0) {
        const int col = frag_col(wn, j, lane) + h;
        red[(wm * BN + col) * 2 + 0] = a;
        red[(wm * BN + col) * 2 + 1] = b;
      }
    }
  asm volatile("bar.sync 1, 256;\n" ::: "memory");
  if (tid < BN) {
    const double s2 = red[tid * 2] + red[(BN + tid) * 2];
    const double s1 = red[tid * 2 + 1] + red[(BN + tid) * 2 + 1];
    const long long c = (long long)blockIdx.x * BN + tid;
    const double v = sf2 - s2;
    var[c] = v > 0.0 ? v : (v != v ? v : 0.0);   // max(.,0), NaN propagates
    mean[c] = mconst + s1;
  }
}

bool g_attr[16] = {false};   // function attributes are per device

}  // namespace

int b7_launch_posterior(b7_ctx* ctx, const double* LinvT, const double* beta, int Np, const double* ksT,
                        int64_t cols_pad, double sf2, double mconst, double* mean, double* var) {
  if (!g_attr[ctx->device & 15]) {
    B7_CUDA(cudaFuncSetAttribute(posterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    g_attr[ctx->device & 15] = true;
  }
  if (cols_pad <= 0) return 0;
  posterior_kernel<<<(unsigned)(cols_pad / BN), P_THREADS, P_SMEM, ctx->stream>>>(LinvT, beta, Np, Np / B7_NB, ksT, sf2,
                                                                               mconst, mean, var);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}
