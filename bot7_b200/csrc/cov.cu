// Covariance builders (kernel #1): K(X,X) (+ noise/jitter on the diagonal) and K(X*,X) tiles,
// ARD squared-exponential and Matern-5/2.
//
// Replaces the kernel evaluation inside gpTorch7's gp_regressor (external; called through
// model:predict at reference scores/expected_improvement.lua:63).  The only in-tree witness of
// that arithmetic is utils.math.pdist (utils/math.lua:65-111), which forms distances in the
// cancellation-prone expanded form |x|^2+|z|^2-2xz through a GEMM and clamps at 0.  Here the
// distance is accumulated directly, r2 = sum_d ((a_d - b_d) w_d)^2 (oracle/SPEC.md), in registers:
// the pass writes 8 B per matrix entry and reads nothing but the (cache-resident) inputs.
//
// Layout: out is rows_pad x Np row-major (k contiguous).  One thread owns one column k (one
// observation) and keeps its d coordinates in registers; the block walks over a chunk of rows
// (candidates) whose coordinates sit in shared memory and are read as warp broadcasts; every
// store instruction writes 32 consecutive doubles.
#include "b7_internal.h"

namespace {

constexpr int kRowsPerBlock = 32;

// exp(x) for x <= 0, table driven: x = n ln2/64 + r, exp(x) = 2^(n>>6) * 2^((n&63)/64) * (1 + p(r)), |r| <= ln2/128,
// p of degree 5 (truncation 3e-17).  ~12 FP64 instructions instead of ~30 for exp(): this kernel is bound by
// the FP64 issue rate, not by the 8 bytes it writes per entry.  <= 1 ulp measured against glibc on 1e7 points;
// results below the normal range (x < -708) are flushed to 0.
__constant__ double c_exp_tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tab) {
  if (!(x > -708.0)) return (x != x) ? x : 0.0;
  const double t = fma(x, 92.33248261689366, 6755399441055744.0);   // 64/ln2, 1.5*2^52: n in the low word
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -0.010830424696249145, x);          // ln2/64 high part
  r = fma(nf, -(3.623510646634843e-19), r);                 // ln2/64 low part
  double p = fma(r, 8.33333333333333322e-03, 4.16666666666666644e-02);
  p = fma(p, r, 1.66666666666666657e-01);
  p = fma(p, r, 0.5);
  p = fma(p * r, r, r);
  const double T = tab[n & 63];
  const double v = fma(T, p, T);
  return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}

template <int DT, int KERNEL>
__global__ void __launch_bounds__(256)
cov_kernel(const double* __restrict__ A, long long rows, long long rows_pad, int d, const double* __restrict__ Xt,
           int N, int Np, const double* __restrict__ par_base, long long par_stride, double* __restrict__ out_base,
           long long out_stride, int is_kxx, int tiled) {
  __shared__ double s_a[kRowsPerBlock][DT];
  __shared__ double s_w[DT];
  __shared__ double s_tab[64];
  if (threadIdx.x < 64) s_tab[threadIdx.x] = c_exp_tab[threadIdx.x];
  const double* par = par_base + (long long)blockIdx.z * par_stride;
  double* out = out_base + (long long)blockIdx.z * out_stride;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * kRowsPerBlock;
  // K(X,X): only 128x128 blocks at or below the diagonal are ever read (lower Cholesky)
  if (is_kxx && (long long)(blockIdx.x * blockDim.x) > (r0 | 127)) return;
  for (int e = threadIdx.x; e < kRowsPerBlock * DT; e += blockDim.x) {
    int r = e / DT, i = e % DT;
    long long row = r0 + r;
    s_a[r][i] = (row < rows && i < d) ? A[row * d + i] : 0.0;
  }
  if (threadIdx.x < DT) s_w[threadIdx.x] = threadIdx.x < d ? par[threadIdx.x] : 0.0;
  __syncthreads();
  if (k >= Np) return;
  const double sf2 = par[B7_MAX_DIMS], diag_add = par[B7_MAX_DIMS + 1];
  double b[DT];
#pragma unroll
  for (int i = 0; i < DT; ++i) b[i] = (i < d && k < N) ? Xt[(long long)i * Np + k] : 0.0;
  const int r_end = (int)((rows_pad - r0) < kRowsPerBlock ? (rows_pad - r0) : kRowsPerBlock);
  for (int r = 0; r < r_end; ++r) {
    const long long row = r0 + r;
    double val;
    if (row >= rows || k >= N) {
      val = (is_kxx && row == k) ? 1.0 : 0.0;   // identity padding keeps the padded factor trivial
    } else {
      double r2 = 0.0;
#pragma unroll
      for (int i = 0; i < DT; ++i) {
        double t = (s_a[r][i] - b[i]) * s_w[i];
        r2 = fma(t, t, r2);
      }
      if (KERNEL == B7_KERNEL_ARDSE) {
        val = sf2 * exp_neg(-0.5 * r2, s_tab);
      } else {
        double rr = sqrt(r2), s5r = 2.23606797749978969641 * rr;
        val = sf2 * ((1.0 + s5r + (5.0 / 3.0) * r2) * exp_neg(-s5r, s_tab));
      }
      if (is_kxx && row == k) val += diag_add;
    }
    if (tiled) {
      // fragment order of gemm_tile.cuh: [row tile][k tile 16][k group 4][row 128][4]
      const long long off = ((row >> 7) * (long long)(Np >> 4) + (k >> 4)) * 2048 + ((k >> 2) & 3) * 512 + (row & 127) * 4 + (k & 3);
      out[off] = val;
    } else {
      out[row * Np + k] = val;
    }
  }
}

template <int DT>
int launch_dt(b7_ctx* ctx, int kernel, const double* A, long long rows, long long rows_pad, int d, const double* Xt,
              int N, int Np, const double* par, long long par_stride, double* out, long long out_stride, int batch,
              bool is_kxx, bool tiled) {
  dim3 grid((Np + 255) / 256, (unsigned)((rows_pad + kRowsPerBlock - 1) / kRowsPerBlock), batch);
  if (kernel == B7_KERNEL_ARDSE)
    cov_kernel<DT, B7_KERNEL_ARDSE><<<grid, 256, 0, ctx->stream>>>(A, rows, rows_pad, d, Xt, N, Np, par, par_stride, out,
                                                                 out_stride, is_kxx, tiled);
  else
    cov_kernel<DT, B7_KERNEL_MATERN52><<<grid, 256, 0, ctx->stream>>>(A, rows, rows_pad, d, Xt, N, Np, par, par_stride,
                                                                    out, out_stride, is_kxx, tiled);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int b7_launch_cov_batched(b7_ctx* ctx, int kernel, const double* A, int64_t rows, int64_t rows_pad, int d,
                          const double* Xt, int N, int Np, const double* par, int64_t par_stride, double* out,
                          int64_t out_stride, int batch, bool is_kxx, bool tiled) {
  if (rows_pad <= 0) return 0;
  if (rows_pad / kRowsPerBlock + 1 > 65535) { b7_set_error("cov: too many rows per launch"); return B7_ERR_ARG; }
  if (d <= 2) return launch_dt<2>(ctx, kernel, A, rows, rows_pad, d, Xt, N, Np, par, par_stride, out, out_stride, batch, is_kxx, tiled);
  if (d <= 4) return launch_dt<4>(ctx, kernel, A, rows, rows_pad, d, Xt, N, Np, par, par_stride, out, out_stride, batch, is_kxx, tiled);
  if (d <= 6) return launch_dt<6>(ctx, kernel, A, rows, rows_pad, d, Xt, N, Np, par, par_stride, out, out_stride, batch, is_kxx, tiled);
  if (d <= 8) return launch_dt<8>(ctx, kernel, A, rows, rows_pad, d, Xt, N, Np, par, par_stride, out, out_stride, batch, is_kxx, tiled);
  if (d <= 16) return launch_dt<16>(ctx, kernel, A, rows, rows_pad, d, Xt, N, Np, par, par_stride, out, out_stride, batch, is_kxx, tiled);
  if (d <= 24) return launch_dt<24>(ctx, kernel, A, rows, rows_pad, d, Xt, N, Np, par, par_stride, out, out_stride, batch, is_kxx, tiled);
  return launch_dt<40>(ctx, kernel, A, rows, rows_pad, d, Xt, N, Np, par, par_stride, out, out_stride, batch, is_kxx, tiled);
}

