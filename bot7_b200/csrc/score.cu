// Fused scoring pass (kernel #3): EI / confidence bound per hyper-parameter draw, sequential
// average over the draws, and first-maximum argmax -- one pass over the candidates.
//
// Replaces (reference) EI.compute scores/expected_improvement.lua:69-88, conf_bound.compute
// scores/confidence_bound.lua:70-106, utils.math.erf / norm_cdf / norm_pdf utils/math.lua:261-312,
// the draw loop of bot:eval bots/bayesopt.lua:69-80 and `score:max(1)` bots/bayesopt.lua:96.
// The reference makes ~14 full-length tensor passes per draw; here each (mean, var) pair is read
// once (16 B per candidate per draw) and 8 B per candidate are written.
//
// Arithmetic follows the reference op for op: every multiply/add is rounded separately
// (__dmul_rn/__dadd_rn keep ptxas from contracting them into FMAs), erf is the reference's
// Abramowitz-Stegun 7.1.26 polynomial, the S-sum starts at 0.0 and runs in draw order with a
// single divide at the end.  IEEE edge cases (sigma = 0, NaN) fall out of the formula.
#include <math.h>

#include "b7_internal.h"

namespace {

struct ScoreConst {
  double sqrt2_inv, sqrt2pi_inv;   // utils/math.lua:13,15 (computed on the host like Lua does)
  double tradeoff, sign, fmin, inv_unused;
  int kind, bound;
};

__device__ __forceinline__ double erf_ref(double x) {
  // utils/math.lua:261-288
  const double c1 = 0.254829592, c2 = -0.284496736, c3 = 1.421413741, c4 = -1.453152027, c5 = 1.061405429,
               p = 0.3275911;
  double t = __drcp_rn(__dadd_rn(__dmul_rn(fabs(x), p), 1.0));   // correctly rounded 1/x == pow(x, -1)
  double r = __dmul_rn(t, c5);
  r = __dadd_rn(r, c4); r = __dmul_rn(r, t);
  r = __dadd_rn(r, c3); r = __dmul_rn(r, t);
  r = __dadd_rn(r, c2); r = __dmul_rn(r, t);
  r = __dadd_rn(r, c1); r = __dmul_rn(r, t);
  double e = exp(__dmul_rn(__dmul_rn(x, x), -1.0));
  r = __dadd_rn(__dmul_rn(__dmul_rn(r, e), -1.0), 1.0);
  double sgn = (x >= 0.0) ? 1.0 : -1.0;   // ge(src,0)*2 - 1
  return __dmul_rn(r, sgn);
}

__device__ __forceinline__ double ei_ref(double mu, double var, const ScoreConst& k) {
  // scores/expected_improvement.lua:73-80
  double sigma = __dsqrt_rn(var);
  double imprv = __dadd_rn(__dadd_rn(k.fmin, -mu), -k.tradeoff);
  double z = __ddiv_rn(imprv, sigma);
  double cdf = __dmul_rn(__dadd_rn(erf_ref(__dmul_rn(z, k.sqrt2_inv)), 1.0), 0.5);   // utils/math.lua:305-312
  double pdf = __dmul_rn(exp(__dmul_rn(__dmul_rn(z, z), -0.5)), k.sqrt2pi_inv);       // utils/math.lua:293-300
  double ei = __dadd_rn(__dmul_rn(imprv, cdf), __dmul_rn(sigma, pdf));
  return ei < 0.0 ? 0.0 : ei;   // clamp(0, inf) is comparison based: NaN passes through
}

__device__ __forceinline__ double cb_ref(double mu, double var, const ScoreConst& k) {
  // scores/confidence_bound.lua:70-106
  double s = __dmul_rn(__dsqrt_rn(var), k.tradeoff);
  double val = (k.bound == B7_BOUND_LOWER) ? __dadd_rn(mu, -s) : __dadd_rn(mu, s);
  return k.sign > 0.0 ? val : -val;
}

__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
  return v > bv || (v == bv && i < bi);
}

__device__ __forceinline__ bool is_removed(const long long* removed, long long n, long long row) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    long long v = removed[mid];
    if (v == row) return true;
    if (v < row) lo = mid + 1; else hi = mid;
  }
  return false;
}

template <int KIND>
__global__ void __launch_bounds__(256)
score_kernel(const double* __restrict__ mean, const double* __restrict__ var, int S, long long M, long long ld,
             ScoreConst k, const long long* __restrict__ removed, long long n_removed, long long row_base,
             double* __restrict__ score_out, double* __restrict__ part_best, long long* __restrict__ part_idx,
             long long* __restrict__ part_nan) {
  double best = -INFINITY;
  long long best_i = LLONG_MAX, nans = 0;
  const double inv_div = (double)S;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < M; c += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;   // bots/bayesopt.lua:70 torch.zeros
    const double* pm = mean + c;
    const double* pv = var + c;
#pragma unroll 4
    for (int s = 0; s < S; ++s) {
      double mu = __ldg(pm + (long long)s * ld), v = __ldg(pv + (long long)s * ld);
      double sc = KIND == B7_SCORE_EI ? ei_ref(mu, v, k) : cb_ref(mu, v, k);
      acc = __dadd_rn(acc, sc);   // bots/bayesopt.lua:76 score:add
    }
    acc = __ddiv_rn(acc, inv_div);   // bots/bayesopt.lua:79 score:div(nSamples)
    long long row = row_base + c;
    bool dead = n_removed > 0 && is_removed(removed, n_removed, row);
    if (score_out) score_out[c] = dead ? __longlong_as_double(0x7ff8000000000000LL) : acc;
    if (!dead) {
      if (acc != acc) ++nans;
      else if (better(acc, row, best, best_i)) { best = acc; best_i = row; }
    }
  }
  // warp-shuffle argmax, then one partial per block
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double ov = __shfl_down_sync(0xffffffffu, best, off);
    long long oi = __shfl_down_sync(0xffffffffu, best_i, off);
    long long on = __shfl_down_sync(0xffffffffu, nans, off);
    if (better(ov, oi, best, best_i)) { best = ov; best_i = oi; }
    nans += on;
  }
  __shared__ double s_v[8];
  __shared__ long long s_i[8], s_n[8];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_v[w] = best; s_i[w] = best_i; s_n[w] = nans; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < (int)(blockDim.x >> 5); ++j) {
      if (better(s_v[j], s_i[j], best, best_i)) { best = s_v[j]; best_i = s_i[j]; }
      nans += s_n[j];
    }
    part_best[blockIdx.x] = best;
    part_idx[blockIdx.x] = best_i;
    part_nan[blockIdx.x] = nans;
  }
}

}  // namespace

int b7_score_grid_size(b7_ctx* ctx, int64_t M) {
  long long want = (M + 255) / 256, cap = (long long)ctx->sm_count * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

int b7_launch_score(b7_ctx* ctx, int kind, const double* mean, const double* var, int S, int64_t M, int64_t ld,
                    double tradeoff, int bound, double sign, double fmin, const int64_t* removed, int64_t n_removed,
                    int64_t row_base, double* score_out, double* part_best, int64_t* part_idx, int64_t* part_nan,
                    int* n_parts) {
  ScoreConst k;
  k.sqrt2_inv = 1.0 / sqrt(2.0);
  k.sqrt2pi_inv = 1.0 / sqrt(2.0 * 3.14159265358979323846);
  k.tradeoff = tradeoff; k.sign = sign; k.fmin = fmin; k.inv_unused = 0; k.kind = kind; k.bound = bound;
  int grid = b7_score_grid_size(ctx, M);
  *n_parts = grid;
  if (kind == B7_SCORE_EI)
    score_kernel<B7_SCORE_EI><<<grid, 256, 0, ctx->stream>>>(mean, var, S, M, ld, k, (const long long*)removed, n_removed,
                                                          row_base, score_out, part_best, (long long*)part_idx,
                                                          (long long*)part_nan);
  else
    score_kernel<B7_SCORE_CB><<<grid, 256, 0, ctx->stream>>>(mean, var, S, M, ld, k, (const long long*)removed, n_removed,
                                                          row_base, score_out, part_best, (long long*)part_idx,
                                                          (long long*)part_nan);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}
