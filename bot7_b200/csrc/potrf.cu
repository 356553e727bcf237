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
//   diag  : one CTA per draw factors the 128x128 diagonal block with the rows held in registers and
//           then inverts it (Gauss-Jordan forward elimination); it also produces x_j = L_jj^-1 r_j,
//           the running log-determinant and info.
//   panel : L21 = A21 * inv(L11)^T as DMMA tiles; the epilogue folds r_i -= L21 x_j, so the forward
//           substitution for beta costs no extra pass over L.
//   trail : A22 -= L21 L21^T on the lower tiles (DMMA).  Two-level blocking: inside an outer panel of
//           4 blocks only the panel's own columns are updated per step (k = 128); the tiles to the
//           right of the panel are updated once per outer panel with k = 512, so most flops run in
//           long-k launches with a quarter of the read-modify-write traffic on C.
// Inversion (LAPACK dtrtri order, in place, column sweep from the right):
//   T = L[j+1:, j] * inv(L_jj)  (stored transposed), then  X[j+1:, j] = -Linv[j+1:, j+1:] * T.
#include "b7_internal.h"
#include "gemm_tile.cuh"

using namespace b7g;

namespace {

constexpr int NBK = B7_NB;          // 128
constexpr int DIAG_THREADS = 512;
// shared: Lcol[128][128] | pivs[128] | invs[128] | rowbuf[2][128] | rj[128] | xpart[4][128] | lg[128]
constexpr int DIAG_SMEM = (NBK * NBK + 11 * NBK) * 8;

// ---- diagonal block: potf2 + inverse + x_j + logdet + info -------------------------------------
// Thread (r, part) keeps 32 entries of row r in registers (static indexing: the column loop is
// unrolled in chunks of 32), so a step is one shared-memory column broadcast, one barrier and
// 32 predicated DFMAs -- no read-modify-write chains through shared memory.  Two sweeps: the
// right-looking Cholesky, then Gauss-Jordan forward elimination for the inverse.
__global__ void __launch_bounds__(DIAG_THREADS, 1)
diag_kernel(double* __restrict__ fac, long long fac_stride, int Np, int j, double* __restrict__ dinv,
            double* __restrict__ dinvT, long long dinv_stride, double* __restrict__ beta, double* __restrict__ logdet,
            int* __restrict__ info, int s0) {
  extern __shared__ double sm[];
  double* Lcol = sm;                       // Lcol[q*128 + r] = A[r][q] as it was when column q was eliminated
  double* pivs = sm + NBK * NBK;
  double* invs = pivs + NBK;
  double* rowbuf = invs + NBK;             // [2][128]
  double* rj = rowbuf + 2 * NBK;
  double* xpart = rj + NBK;                // [4][128]
  double* lg = xpart + 4 * NBK;
  const int s = s0 + blockIdx.x, tid = threadIdx.x;
  const int r = tid & 127, part = tid >> 7, c0 = part * 32;
  double* blk = fac + (long long)s * fac_stride + (long long)j * NBK * Np + (long long)j * NBK;
  double a[32];
  {
    const double2* src = reinterpret_cast<const double2*>(blk + (long long)r * Np + c0);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double2 v = src[kk];
      a[2 * kk] = v.x;
      a[2 * kk + 1] = v.y;
    }
  }
  if (tid < NBK) rj[tid] = beta[(long long)s * Np + j * NBK + tid];
  int my_info = 0;
  // ---- sweep 1: Cholesky ----
  for (int pq = 0; pq < 4; ++pq) {
#pragma unroll
    for (int qq = 0; qq < 32; ++qq) {
      const int q = pq * 32 + qq;
      if (part == pq) Lcol[q * NBK + r] = a[qq];
      __syncthreads();
      const double piv = Lcol[q * NBK + q];
      const double inv = rsqrt(piv);
      if (tid == 0) {
        pivs[q] = piv;
        invs[q] = inv;
        if (!(piv > 0.0) && my_info == 0) my_info = j * NBK + q + 1;
      }
      const double arq = Lcol[q * NBK + r];
      if (c0 + 31 > q) {   // warp-uniform: this 32-column slice still has columns right of q
        const double t = arq * (inv * inv);   // A_rq / pivot
        const double2* colv = reinterpret_cast<const double2*>(Lcol + q * NBK + c0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double2 cv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) cv[u] = colv[h * 8 + u];   // broadcast loads, issued together
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int kk = h * 16 + 2 * u, k = c0 + kk;
            const double n0 = fma(-t, cv[u].x, a[kk]), n1 = fma(-t, cv[u].y, a[kk + 1]);
            a[kk] = (k > q && k <= r) ? n0 : a[kk];
            a[kk + 1] = (k + 1 > q && k + 1 <= r) ? n1 : a[kk + 1];
          }
        }
      }
      if (part == pq) {
        if (r > q) a[qq] = arq * inv;
        else if (r == q) a[qq] = piv * inv;
      }
    }
  }
  // L block back to global (upper part zeroed)
  {
    double2* dst = reinterpret_cast<double2*>(blk + (long long)r * Np + c0);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const int k = c0 + 2 * kk;
      dst[kk] = make_double2(k <= r ? a[2 * kk] : 0.0, k + 1 <= r ? a[2 * kk + 1] : 0.0);
    }
  }
  __syncthreads();   // invs complete
  // ---- sweep 2: inverse by forward elimination, W starts as the identity ----
  double w[32];
#pragma unroll
  for (int kk = 0; kk < 32; ++kk) w[kk] = (c0 + kk == r) ? 1.0 : 0.0;
  for (int pq = 0; pq < 4; ++pq) {
#pragma unroll
    for (int qq = 0; qq < 32; ++qq) {
      const int q = pq * 32 + qq;
      const double inv = invs[q];
      double* buf = rowbuf + (q & 1) * NBK;
      if (r == q) {
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
          w[kk] *= inv;            // X[q][c] = W[q][c] / l_qq
          buf[c0 + kk] = w[kk];
        }
      }
      __syncthreads();
      if (c0 <= q) {   // warp-uniform: this slice has columns <= q
        const double liq = Lcol[q * NBK + r] * inv;   // L[r][q]
        const double2* rowv = reinterpret_cast<const double2*>(buf + c0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double2 rv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) rv[u] = rowv[h * 8 + u];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int kk = h * 16 + 2 * u, c = c0 + kk;
            const double n0 = fma(-liq, rv[u].x, w[kk]), n1 = fma(-liq, rv[u].y, w[kk + 1]);
            w[kk] = (r > q && c <= q) ? n0 : w[kk];
            w[kk + 1] = (r > q && c + 1 <= q) ? n1 : w[kk + 1];
          }
        }
      }
    }
  }
  // x_j = inv(L_jj) r_j ; outputs
  double xp = 0.0;
#pragma unroll
  for (int kk = 0; kk < 32; ++kk)
    if (c0 + kk <= r) xp = fma(w[kk], rj[c0 + kk], xp);
  xpart[part * NBK + r] = xp;
  if (tid < NBK) lg[tid] = log(pivs[tid] * invs[tid]);
  double* di = dinv + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  double* dt = dinvT + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  {
    double2* dst = reinterpret_cast<double2*>(di + r * NBK + c0);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const int k = c0 + 2 * kk;
      dst[kk] = make_double2(k <= r ? w[2 * kk] : 0.0, k + 1 <= r ? w[2 * kk + 1] : 0.0);
    }
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) dt[(c0 + kk) * NBK + r] = (c0 + kk <= r) ? w[kk] : 0.0;
  }
  __syncthreads();
  if (tid < NBK)
    beta[(long long)s * Np + j * NBK + tid] = ((xpart[tid] + xpart[NBK + tid]) + xpart[2 * NBK + tid]) + xpart[3 * NBK + tid];
  if (tid == 0) {
    double ld_acc = 0.0;
    for (int q = 0; q < NBK; ++q) ld_acc += lg[q];
    logdet[s] = (j == 0 ? 0.0 : logdet[s]) + ld_acc;
    if (j == 0) info[s] = my_info;
    else if (info[s] == 0 && my_info != 0) info[s] = my_info;
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

// ---- trailing update: C[it][nt] -= L[it][kb0:kb1] L[nt][kb0:kb1]^T on lower tiles ---------------
// it = it0 + blockIdx.x, nt = nt0 + blockIdx.y (tiles above the diagonal exit at once).  Used with
// one k block inside the current outer panel and with the whole outer panel (k = 128 * W) for the
// tiles to its right: the long-k launches carry most of the flops with one read-modify-write of C.
__global__ void __launch_bounds__(THREADS, 1)
trail_kernel(double* __restrict__ fac, long long fac_stride, int Np, int kb0, int kb1, int it0, int nt0, int s0) {
  extern __shared__ __align__(16) double smem[];
  const int it = it0 + blockIdx.x, nt = nt0 + blockIdx.y;
  if (nt > it) return;
  const int s = s0 + blockIdx.z;
  double* base = fac + (long long)s * fac_stride;
  const double* A = base + (long long)it * NBK * Np + (long long)kb0 * NBK;
  const double* B = base + (long long)nt * NBK * Np + (long long)kb0 * NBK;
  double* C = base + (long long)it * NBK * Np + (long long)nt * NBK;
  Acc acc; acc.zero();
  mainloop(A, Np, B, Np, (kb1 - kb0) * (NBK / BK), smem, acc);
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
  const int W = 4;   // outer panel = 4 blocks (512 columns)
  for (int J = 0; J < NB; J += W) {
    const int Jend = J + W < NB ? J + W : NB;
    for (int j = J; j < Jend; ++j) {
      diag_kernel<<<count, DIAG_THREADS, DIAG_SMEM, ctx->stream>>>(gp->fac, fs, Np, j, gp->dinv, gp->dinvT, ds, gp->beta,
                                                                  gp->logdet, gp->info, s0);
      b7_count(ctx);
      const int rem = NB - 1 - j;
      if (rem > 0) {
        panel_kernel<<<dim3(rem, 1, count), THREADS, SMEM_BYTES, ctx->stream>>>(gp->fac, fs, Np, j, gp->dinv, ds, gp->beta, s0);
        b7_count(ctx);
      }
      if (Jend - 1 - j > 0) {   // columns j+1 .. Jend-1 of the outer panel, all rows below
        trail_kernel<<<dim3(rem, Jend - 1 - j, count), THREADS, SMEM_BYTES, ctx->stream>>>(gp->fac, fs, Np, j, j + 1, j + 1, j + 1, s0);
        b7_count(ctx);
      }
    }
    const int right = NB - Jend;
    if (right > 0) {
      trail_kernel<<<dim3(right, right, count), THREADS, SMEM_BYTES, ctx->stream>>>(gp->fac, fs, Np, J, Jend, Jend, Jend, s0);
      b7_count(ctx);
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
