// Batched fp64 blocked Cholesky (kernel #2) with beta = L^-1 (y - m), log-determinant and LAPACK
// style info, plus the in-place triangular inversion used by the posterior pass.
//
// Replaces what Torch7 reaches through torch.potrf in utils.math.chol (reference
// utils/math.lua:159-218; LAPACK dpotrf, one matrix at a time) and the two triangular solves of
// gpTorch7's posterior.  One factor per slice-sampled hyper-parameter draw, all draws in one launch
// sequence (grid.z / grid.x = draw).
//
// Right-looking, block size 128 (= the DMMA tile edge of gemm_tile.cuh), row-major lower storage,
// matrices padded to a multiple of 128 with identity.  Per block column j:
//   diag  : one CTA per draw factors the 128x128 diagonal block in shared memory and, fused in the
//           same sweep, inverts it (Gauss-Jordan on the lower triangle; the inverse lives in the
//           upper triangle of the same shared array); it also produces x_j = L_jj^-1 r_j, the running
//           log-determinant and info.
//   panel : L21 = A21 * inv(L11)^T as DMMA tiles; the epilogue folds r_i -= L21 x_j, so the forward
//           substitution for beta costs no extra pass over L.
//   trail : A22 -= L21 L21^T on the lower tiles (DMMA).
// Inversion (LAPACK dtrtri order, in place, column sweep from the right):
//   T = L[j+1:, j] * inv(L_jj)  (stored transposed), then  X[j+1:, j] = -Linv[j+1:, j+1:] * T.
#include "b7_internal.h"
#include "gemm_tile.cuh"

using namespace b7g;

namespace {

constexpr int NBK = B7_NB;          // 128
constexpr int DLD = NBK + 1;        // leading dimension of the diag workspace (129 doubles)
constexpr int DIAG_THREADS = 512;
constexpr int DIAG_SMEM = NBK * DLD * 8 + NBK * 8;

// ---- diagonal block: potf2 + inverse + x_j + logdet + info -------------------------------------
__global__ void __launch_bounds__(DIAG_THREADS)
diag_kernel(double* __restrict__ fac, long long fac_stride, int Np, int j, double* __restrict__ dinv,
            double* __restrict__ dinvT, long long dinv_stride, double* __restrict__ beta, double* __restrict__ logdet,
            int* __restrict__ info, int s0) {
  extern __shared__ double sm[];
  double* a = sm;                   // a[i*DLD + k]: k<=i lower of A/L ; k>i: W/X[k-1][i] (inverse, transposed)
  double* xr = sm + NBK * DLD;      // r_j, then x_j
  const int s = s0 + blockIdx.x, tid = threadIdx.x;
  double* blk = fac + (long long)s * fac_stride + (long long)j * NBK * Np + (long long)j * NBK;
  for (int e = tid; e < NBK * NBK; e += DIAG_THREADS) {
    int i = e >> 7, k = e & 127;
    double v = blk[(long long)i * Np + k];
    if (k <= i) a[i * DLD + k] = v;
    if (k > i) a[i * DLD + k + 1] = 0.0;     // W[k][i] = 0 for i < k
    if (k == i) a[i * DLD + i + 1] = 1.0;    // W[i][i] = 1
  }
  if (tid < NBK) xr[tid] = beta[(long long)s * Np + j * NBK + tid];
  double ld_acc = 0.0;
  int my_info = 0;
  __syncthreads();
  for (int q = 0; q < NBK; ++q) {
    const double piv = a[q * DLD + q];
    const double lq = sqrt(piv);
    const double inv = 1.0 / lq;
    if (tid == 0) {
      if (!(piv > 0.0) && my_info == 0) my_info = j * NBK + q + 1;
      ld_acc += log(lq);
    }
    __syncthreads();
    // scale column q of L (rows > q) and row q of the inverse (cols <= q)
    if (tid < NBK) {
      if (tid > q) a[tid * DLD + q] *= inv;
      else a[tid * DLD + q + 1] *= inv;      // X[q][tid] = W[q][tid] / l_qq
      if (tid == q) a[q * DLD + q] = lq;
    }
    __syncthreads();
    // rows i > q: trailing update (k' > q) and inverse update (k' <= q)
    const int n_rows = NBK - 1 - q;
    for (int e = tid; e < n_rows * NBK; e += DIAG_THREADS) {
      const int i = q + 1 + (e >> 7), kk = e & 127;
      if (kk > i) continue;
      const double liq = a[i * DLD + q];
      if (kk > q) a[i * DLD + kk] -= liq * a[kk * DLD + q];
      else a[kk * DLD + i + 1] -= liq * a[kk * DLD + q + 1];
    }
    __syncthreads();
  }
  // x_j = inv(L_jj) r_j
  double xv = 0.0;
  if (tid < NBK) {
    for (int c = 0; c <= tid; ++c) xv += a[c * DLD + tid + 1] * xr[c];
  }
  __syncthreads();
  if (tid < NBK) beta[(long long)s * Np + j * NBK + tid] = xv;
  if (tid == 0) {
    logdet[s] = (j == 0 ? 0.0 : logdet[s]) + ld_acc;
    if (j == 0) info[s] = my_info;
    else if (info[s] == 0 && my_info != 0) info[s] = my_info;
  }
  double* di = dinv + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  double* dt = dinvT + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  for (int e = tid; e < NBK * NBK; e += DIAG_THREADS) {
    int i = e >> 7, k = e & 127;
    blk[(long long)i * Np + k] = (k <= i) ? a[i * DLD + k] : 0.0;
    di[e] = (k <= i) ? a[k * DLD + i + 1] : 0.0;   // X[i][k]
    dt[e] = (k >= i) ? a[i * DLD + k + 1] : 0.0;   // X[k][i]
  }
}

// ---- panel: L21 = A21 * inv(L11)^T, beta_i -= L21 x_j --------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
panel_kernel(double* __restrict__ fac, long long fac_stride, int Np, int j, const double* __restrict__ dinv,
             long long dinv_stride, double* __restrict__ beta, int s0) {
  extern __shared__ __align__(16) double smem[];
  const int s = s0 + blockIdx.z, it = j + 1 + blockIdx.x;
  double* tile = fac + (long long)s * fac_stride + (long long)it * NBK * Np + (long long)j * NBK;
  const double* B = dinv + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  Acc acc; acc.zero();
  mainloop(tile, Np, B, NBK, NBK / BK, smem, acc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 2, wn = warp & 3;
  const double* xj = beta + (long long)s * Np + j * NBK;
  double* red = smem;   // [128][4]
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = frag_row(wm, i, lane);
    double part = 0.0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int col = frag_col(wn, jj, lane);
      *reinterpret_cast<double2*>(tile + (long long)row * Np + col) = make_double2(acc.c[i][jj][0], acc.c[i][jj][1]);
      part += acc.c[i][jj][0] * xj[col] + acc.c[i][jj][1] * xj[col + 1];
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if ((lane & 3) == 0) red[row * 4 + wn] = part;
  }
  __syncthreads();
  if (tid < NBK) {
    double sum = ((red[tid * 4 + 0] + red[tid * 4 + 1]) + red[tid * 4 + 2]) + red[tid * 4 + 3];
    beta[(long long)s * Np + it * NBK + tid] -= sum;
  }
}

// ---- trailing update: A22 -= L21 L21^T (lower tiles) ---------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
trail_kernel(double* __restrict__ fac, long long fac_stride, int Np, int j, int s0) {
  extern __shared__ __align__(16) double smem[];
  const int it = j + 1 + blockIdx.x, nt = j + 1 + blockIdx.y;
  if (nt > it) return;
  const int s = s0 + blockIdx.z;
  double* base = fac + (long long)s * fac_stride;
  const double* A = base + (long long)it * NBK * Np + (long long)j * NBK;
  const double* B = base + (long long)nt * NBK * Np + (long long)j * NBK;
  double* C = base + (long long)it * NBK * Np + (long long)nt * NBK;
  Acc acc; acc.zero();
  mainloop(A, Np, B, Np, NBK / BK, smem, acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      double2* p = reinterpret_cast<double2*>(C + (long long)frag_row(wm, i, lane) * Np + frag_col(wn, jj, lane));
      double2 v = *p;
      v.x -= acc.c[i][jj][0];
      v.y -= acc.c[i][jj][1];
      *p = v;
    }
}

// ---- inversion sweep ------------------------------------------------------------------------------
__global__ void place_diag_kernel(double* __restrict__ fac, long long fac_stride, int Np, const double* __restrict__ dinv,
                                  long long dinv_stride, int s0) {
  const int s = s0 + blockIdx.y, j = blockIdx.x;
  double* blk = fac + (long long)s * fac_stride + (long long)j * NBK * Np + (long long)j * NBK;
  const double* di = dinv + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  for (int e = threadIdx.x; e < NBK * NBK; e += blockDim.x) blk[(long long)(e >> 7) * Np + (e & 127)] = di[e];
}

// T = L[it, j] * inv(L_jj), written transposed: tt[c][it*128 + row]
__global__ void __launch_bounds__(THREADS, 1)
inv_step1_kernel(const double* __restrict__ fac, long long fac_stride, int Np, int j, const double* __restrict__ dinvT,
                 long long dinv_stride, double* __restrict__ tt, long long tt_stride, int s0) {
  extern __shared__ __align__(16) double smem[];
  const int s = s0 + blockIdx.z, it = j + 1 + blockIdx.x;
  const double* A = fac + (long long)s * fac_stride + (long long)it * NBK * Np + (long long)j * NBK;
  const double* B = dinvT + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  Acc acc; acc.zero();
  mainloop(A, Np, B, NBK, NBK / BK, smem, acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
  double* T = tt + (long long)s * tt_stride + (long long)it * NBK;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int row = frag_row(wm, i, lane), col = frag_col(wn, jj, lane);
      T[(long long)col * Np + row] = acc.c[i][jj][0];
      T[(long long)(col + 1) * Np + row] = acc.c[i][jj][1];
    }
}

// X[it, j] = - sum_{k=(j+1)*128}^{(it+1)*128-1} Linv[it, k] * T[k, :]
__global__ void __launch_bounds__(THREADS, 1)
inv_step2_kernel(double* __restrict__ fac, long long fac_stride, int Np, int j, const double* __restrict__ tt,
                 long long tt_stride, int s0) {
  extern __shared__ __align__(16) double smem[];
  // longest rows first: better tail behaviour
  const int n_it = gridDim.x, it = j + n_it - (int)blockIdx.x;
  const int s = s0 + blockIdx.z;
  double* base = fac + (long long)s * fac_stride;
  const long long k0 = (long long)(j + 1) * NBK;
  const double* A = base + (long long)it * NBK * Np + k0;
  const double* B = tt + (long long)s * tt_stride + k0;
  double* C = base + (long long)it * NBK * Np + (long long)j * NBK;
  Acc acc; acc.zero();
  mainloop(A, Np, B, Np, (it - j) * (NBK / BK), smem, acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      *reinterpret_cast<double2*>(C + (long long)frag_row(wm, i, lane) * Np + frag_col(wn, jj, lane)) =
          make_double2(-acc.c[i][jj][0], -acc.c[i][jj][1]);
}

bool g_attr_done = false;
int set_attrs() {
  if (g_attr_done) return 0;
  B7_CUDA(cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
  B7_CUDA(cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  B7_CUDA(cudaFuncSetAttribute(trail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  B7_CUDA(cudaFuncSetAttribute(inv_step1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  B7_CUDA(cudaFuncSetAttribute(inv_step2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  g_attr_done = true;
  return 0;
}

}  // namespace

int b7_launch_potrf(b7_gp* gp, int s0, int count) {
  b7_ctx* ctx = gp->ctx;
  B7_CHECK(set_attrs());
  const int Np = gp->Np, NB = gp->NB;
  const long long fs = (long long)Np * Np, ds = (long long)NB * NBK * NBK;
  for (int j = 0; j < NB; ++j) {
    diag_kernel<<<count, DIAG_THREADS, DIAG_SMEM, ctx->stream>>>(gp->fac, fs, Np, j, gp->dinv, gp->dinvT, ds, gp->beta,
                                                                gp->logdet, gp->info, s0);
    b7_count(ctx);
    const int rem = NB - 1 - j;
    if (rem > 0) {
      panel_kernel<<<dim3(rem, 1, count), THREADS, SMEM_BYTES, ctx->stream>>>(gp->fac, fs, Np, j, gp->dinv, ds, gp->beta, s0);
      trail_kernel<<<dim3(rem, rem, count), THREADS, SMEM_BYTES, ctx->stream>>>(gp->fac, fs, Np, j, s0);
      b7_count(ctx, 2);
    }
  }
  B7_CUDA(cudaGetLastError());
  return 0;
}

int b7_launch_trtri(b7_gp* gp, int s0, int count) {
  b7_ctx* ctx = gp->ctx;
  B7_CHECK(set_attrs());
  const int Np = gp->Np, NB = gp->NB;
  const long long fs = (long long)Np * Np, ds = (long long)NB * NBK * NBK, ts = (long long)NBK * Np;
  place_diag_kernel<<<dim3(NB, count), 256, 0, ctx->stream>>>(gp->fac, fs, Np, gp->dinv, ds, s0);
  b7_count(ctx);
  for (int j = NB - 2; j >= 0; --j) {
    const int rem = NB - 1 - j;
    inv_step1_kernel<<<dim3(rem, 1, count), THREADS, SMEM_BYTES, ctx->stream>>>(gp->fac, fs, Np, j, gp->dinvT, ds, gp->tt, ts, s0);
    inv_step2_kernel<<<dim3(rem, 1, count), THREADS, SMEM_BYTES, ctx->stream>>>(gp->fac, fs, Np, j, gp->tt, ts, s0);
    b7_count(ctx, 2);
  }
  B7_CUDA(cudaGetLastError());
  return 0;
}
