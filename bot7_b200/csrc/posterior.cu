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
// sum(v^2) and sum(v*beta) held in registers, and moves on.  V is never written; K* (panel x Np,
// k contiguous) is the only per-candidate operand and is streamed from L2/HBM.  The cp.async
// pipeline runs across row-block boundaries without draining.  All CTAs execute the same
// schedule, so the L^-1 tiles they share hit in L2 and a grid of 148 CTAs is one full wave with no
// tail.  Summation order is fixed (row blocks ascending, fixed shuffle tree): run-to-run
// deterministic.
#include "b7_internal.h"
#include "gemm_tile.cuh"

using namespace b7g;

namespace {

__global__ void __launch_bounds__(THREADS, 1)
posterior_kernel(const double* __restrict__ Linv, const double* __restrict__ beta, int Np, int NB,
                 const double* __restrict__ ks, double sf2, double mconst, double* __restrict__ mean,
                 double* __restrict__ var) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 2, wn = warp & 3;
  const double* gB = ks + (long long)blockIdx.x * BN * Np;
  const int KPB = B7_NB / BK;   // k-tiles per 128-block
  const long long total = (long long)KPB * NB * (NB + 1) / 2;

  // loader cursor
  int l_rb = 0, l_kt = 0;
  auto issue = [&](int slot) {
    double* st = smem + slot * STAGE_DOUBLES;
    load_operand(st, Linv + (long long)l_rb * B7_NB * Np + (long long)l_kt * BK, Np, tid);
    load_operand(st + OPERAND_DOUBLES, gB + (long long)l_kt * BK, Np, tid);
    if (++l_kt == (l_rb + 1) * KPB) { l_kt = 0; ++l_rb; }
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < total) issue(s);
    cp_commit();
  }
  Acc acc; acc.zero();
  double sum2[4][2], sum1[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) sum2[j][0] = sum2[j][1] = sum1[j][0] = sum1[j][1] = 0.0;
  int c_rb = 0, c_kt = 0;
  for (long long it = 0; it < total; ++it) {
    cp_wait<STAGES - 2>();
    __syncthreads();
    if (it + STAGES - 1 < total) issue((int)((it + STAGES - 1) % STAGES));
    cp_commit();
    const double* st = smem + (it % STAGES) * STAGE_DOUBLES;
    compute_stage(st, st + OPERAND_DOUBLES, wm, wn, lane, acc);
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
  cp_wait<0>();
  __syncthreads();
  // reduce over the 8 row lanes of the warp (lane >> 2), then over the two row warps (wm)
  double* red = smem;   // [2][128][2]
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
  __syncthreads();
  if (tid < BN) {
    const double s2 = red[tid * 2] + red[(BN + tid) * 2];
    const double s1 = red[tid * 2 + 1] + red[(BN + tid) * 2 + 1];
    const long long c = (long long)blockIdx.x * BN + tid;
    const double v = sf2 - s2;
    var[c] = v > 0.0 ? v : (v != v ? v : 0.0);   // max(.,0), NaN propagates
    mean[c] = mconst + s1;
  }
}

bool g_attr = false;

}  // namespace

int b7_launch_posterior(b7_ctx* ctx, const double* Linv, const double* beta, int Np, const double* ks,
                        int64_t cols_pad, double sf2, double mconst, double* mean, double* var) {
  if (!g_attr) {
    B7_CUDA(cudaFuncSetAttribute(posterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    g_attr = true;
  }
  if (cols_pad <= 0) return 0;
  posterior_kernel<<<(unsigned)(cols_pad / BN), THREADS, SMEM_BYTES, ctx->stream>>>(Linv, beta, Np, Np / B7_NB, ks, sf2,
                                                                                 mconst, mean, var);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}
