// Batched fp64 blocked Cholesky (kernel #2) with beta = L^-1 (y - m), log-determinant and LAPACK
// style info, plus the in-place triangular inversion used by the posterior pass.
//
// Replaces what Torch7 reaches through torch.potrf in utils.math.chol (reference
// utils/math.lua:159-218; LAPACK dpotrf, one matrix at a time) and the two triangular solves of
// gpTorch7's posterior.  One factor per slice-sampled hyper-parameter draw, all draws in one launch
// sequence (grid.z / grid.x = draw).
//
// Right-looking, block size 128 (= the DMMA tile edge of gemm_tile.cuh).  The matrices live in HBM in
// the tiled (fragment-order) layout of gemm_tile.cuh from the moment K is built until L^-1 is read by
// the posterior pass, so every operand k-step of every GEMM here is one 16 KB TMA bulk copy; lower
// storage, padded to a multiple of 128 with identity.  Per block column j:
//   diag  : one CTA per draw factors and inverts the 128x128 diagonal block inside shared memory,
//           blocked by 16 so that the work is DMMA fragment products; it also produces
//           x_j = L_jj^-1 r_j, the running log-determinant and info.
//   panel : L21 = A21 * inv(L11)^T as DMMA tiles; the epilogue folds r_i -= L21 x_j, so the forward
//           substitution for beta costs no extra pass over L.
//   trail : A22 -= L21 L21^T on the lower tiles (DMMA).  Two-level blocking: inside an outer panel of
//           4 blocks only the panel's own columns are updated per step (k = 128); the tiles to the
//           right of the panel are updated once per outer panel with k = 512, so most flops run in
//           long-k launches with a quarter of the read-modify-write traffic on C.
// Inversion (LAPACK dtrtri order, in place, column sweep from the right):
//   T = L[j+1:, j] * inv(L_jj)  (stored transposed), then  X[j+1:, j] = -Linv[j+1:, j+1:] * T.
#include <stdlib.h>

#include "b7_internal.h"
#include "gemm_tile.cuh"

using namespace b7g;

namespace {

constexpr int NBK = B7_NB;          // 128
constexpr int DIAG_THREADS = 512;
constexpr int DIAG_WARPS = DIAG_THREADS / 32;
constexpr int DLD = 132;            // 132 = 4 mod 16: DMMA fragment loads (row- and column-wise) are conflict free
constexpr int SB = 16;              // sub-block width inside the 128 block
constexpr int DSLD = 20;            // leading dimension of the 16x16 sub-block inverses
constexpr int TLD = 68;             // scratch of the recursive inversion (64 x 68)
// shared: M[128][132] | DS[8][16][20] | T[64][68] | LD[16][17] | pivs,invs,rj,lg [4][128]
constexpr int DIAG_SMEM = (NBK * DLD + 8 * SB * DSLD + 64 * TLD + SB * 17 + 4 * NBK) * 8;

// one 8x8 accumulator fragment: c += sign * A[m.., k0..k1) * op(B);  A row-major (k contiguous).
// NN = false: B is [n][k] row-major (C = A B^T);  NN = true: B is [k][n] row-major (C = A B).
template <bool NN>
__device__ __forceinline__ void frag_mac(double& c0, double& c1, const double* __restrict__ A, int lda,
                                         const double* __restrict__ B, int ldb, int k_begin, int k_end, double sign, int lane) {
  const int r = lane >> 2, c = lane & 3;
  for (int k = k_begin; k < k_end; k += 4) {
    const double a = sign * A[r * lda + k + c];
    const double b = NN ? B[(k + c) * ldb + r] : B[r * ldb + k + c];
    dmma884(c0, c1, a, b);
  }
}

// 1/sqrt(x) without the library's out-of-line slow path (its CALL forces the 16 row registers of the
// factorisation below through local memory): float seed + three Newton steps, <= 2 ulp for normal x.
// A normal positive x is first scaled by an even power of two into [1, 4) (exact, on the exponent bits), so that a
// pivot above FLT_MAX or below FLT_MIN no longer loses the float seed (inf -> seed 0 -> result 0; denormal -> NaN);
// x <= 0, NaN, inf or denormal keeps the plain path and gives inf / NaN / 0, which the info logic flags.
__device__ __forceinline__ double rsqrt_nr(double x) {
  const long long bits = __double_as_longlong(x);
  const int ef = (int)((bits >> 52) & 0x7ff);
  double m = x, scale = 1.0;
  if (bits > 0 && ef >= 1 && ef <= 2046) {
    const int e2 = (ef - 1023) & ~1;                                        // even, rounds towards -inf
    m = __longlong_as_double(bits - ((long long)e2 << 52));                 // x 2^-e2 in [1, 4)
    scale = __longlong_as_double((long long)(1023 - e2 / 2) << 52);         // 2^(-e2 / 2)
  }
  double y = (double)rsqrtf((float)m);
  const double hx = 0.5 * m;
#pragma unroll
  for (int it = 0; it < 3; ++it) y = y * fma(-hx * y, y, 1.5);
  return y * scale;
}

// ---- diagonal block: potf2 + inverse + x_j + logdet + info -------------------------------------
// Blocked inside shared memory so that almost all arithmetic is DMMA on 8x8 fragments:
//   for each 16-wide sub-block: warp 0 factors the 16x16 diagonal piece in registers (shuffles) and
//   inverts it; all warps then form the sub-panel L21 = A21 inv(L11)^T and the trailing update
//   A22 -= L21 L21^T as fragment products.  The 128x128 inverse is assembled afterwards by
//   recursive doubling ([[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]]) for h = 16, 32, 64,
//   again as fragment products, in place.
__global__ void __launch_bounds__(DIAG_THREADS, 1)
diag_kernel(double* __restrict__ fac, long long fac_stride, int Np, int j, double* __restrict__ dinv,
            double* __restrict__ dinvT, long long dinv_stride, double* __restrict__ beta, double* __restrict__ logdet,
            int* __restrict__ info, int s0) {
  extern __shared__ __align__(16) double sm[];
  double* M = sm;
  double* DS = M + NBK * DLD;
  double* T = DS + 8 * SB * DSLD;
  double* LD = T + 64 * TLD;
  double* pivs = LD + SB * 17;
  double* invs = pivs + NBK;
  double* rj = invs + NBK;
  double* lg = rj + NBK;
  const int s = s0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = lane >> 2, c = lane & 3;
  double* blk = fac + (long long)s * fac_stride + tile_off(Np / TILE_K, j, j * (NBK / TILE_K));   // 8 tiles of block (j, j)
  for (int e = tid; e < NBK * NBK; e += DIAG_THREADS) {
    // e enumerates the tiled order [kt][g4][row][kk]
    const int kk = e & 3, i = (e >> 2) & 127, k = (e >> 11) * TILE_K + ((e >> 9) & 3) * 4 + kk;
    M[i * DLD + k] = (k <= i) ? blk[e] : 0.0;
  }
  if (tid < NBK) rj[tid] = beta[(long long)s * Np + j * NBK + tid];
  int my_info = 0;
  __syncthreads();

  for (int p = 0; p < NBK / SB; ++p) {
    const int j0 = p * SB, i0 = j0 + SB, G = (NBK - i0) / 8;
    // (a) 16x16 diagonal piece: factor (lane = row, column broadcast by shuffle), then invert by columns
    if (warp == 0) {
      double a[SB];
#pragma unroll
      for (int k = 0; k < SB; ++k) a[k] = (lane < SB && k <= lane) ? M[(j0 + lane) * DLD + j0 + k] : 0.0;
#pragma unroll
      for (int q = 0; q < SB; ++q) {
        const double piv = __shfl_sync(0xffffffffu, a[q], q);
        const double inv = rsqrt_nr(piv);
        if (lane == 0) {
          pivs[j0 + q] = piv;
          invs[j0 + q] = inv;
          // LAPACK info: first pivot that is not positive -- or whose reciprocal root is not a finite positive number
          if ((!(piv > 0.0) || !(inv > 0.0) || inv > 1.79e308) && my_info == 0) my_info = j * NBK + j0 + q + 1;
        }
        const double my = a[q] * inv;   // L[lane][q] for lane > q
#pragma unroll
        for (int k = 1; k < SB; ++k) {      // constant trip count: keeps a[] in registers
          if (k > q) {
            const double lk = __shfl_sync(0xffffffffu, my, k);
            if (lane >= k) a[k] = fma(-my, lk, a[k]);
          }
        }
        a[q] = (lane > q) ? my : (lane == q ? piv * inv : a[q]);
      }
      if (lane < SB) {
#pragma unroll
        for (int k = 0; k < SB; ++k) {
          const double v = (k <= lane) ? a[k] : 0.0;
          M[(j0 + lane) * DLD + j0 + k] = v;
          LD[lane * 17 + k] = v;
        }
      }
      __syncwarp();
      // column `lane` of the inverse: x_i = (delta_ic - sum_{k<i} L_ik x_k) / L_ii   (x_k = 0 for k < c)
      double x[SB];
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        double acc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < SB; ++k)
          if (k < i) acc = fma(-LD[i * 17 + k], x[k], acc);
        x[i] = (i >= lane) ? acc * invs[j0 + i] : 0.0;
      }
      if (lane < SB) {
#pragma unroll
        for (int i = 0; i < SB; ++i) DS[(p * SB + i) * DSLD + lane] = x[i];
      }
    }
    __syncthreads();
    // (b) sub-panel: rows below, L21 = A21 * inv(L11)^T; one warp owns all 16 columns of its 8 rows
    for (int g = warp; g < G; g += DIAG_WARPS) {
      double* Arow = M + (i0 + 8 * g) * DLD + j0;
      double a4[4];
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) a4[k4] = Arow[r * DLD + 4 * k4 + c];
      double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        dmma884(c00, c01, a4[k4], DS[(p * SB + r) * DSLD + 4 * k4 + c]);
        dmma884(c10, c11, a4[k4], DS[(p * SB + 8 + r) * DSLD + 4 * k4 + c]);
      }
      __syncwarp();
      *reinterpret_cast<double2*>(Arow + r * DLD + 2 * c) = make_double2(c00, c01);
      *reinterpret_cast<double2*>(Arow + r * DLD + 8 + 2 * c) = make_double2(c10, c11);
    }
    __syncthreads();
    // (c) trailing update on the lower fragments: A22 -= L21 L21^T
    const int n_frag = G * (G + 1) / 2;
    for (int f = warp; f < n_frag; f += DIAG_WARPS) {
      int gi = 0, rem = f;
      while (rem > gi) { rem -= gi + 1; ++gi; }
      const int gk = rem;
      double2* cp = reinterpret_cast<double2*>(M + (i0 + 8 * gi + r) * DLD + i0 + 8 * gk + 2 * c);
      double2 cv = *cp;
      frag_mac<false>(cv.x, cv.y, M + (i0 + 8 * gi) * DLD + j0, DLD, M + (i0 + 8 * gk) * DLD + j0, DLD, 0, SB, -1.0, lane);
      *cp = cv;
    }
    __syncthreads();
  }
  // L block back to global (upper part zero)
  for (int e = tid; e < NBK * NBK; e += DIAG_THREADS) {
    const int kk = e & 3, i = (e >> 2) & 127, k = (e >> 11) * TILE_K + ((e >> 9) & 3) * 4 + kk;
    blk[e] = (k <= i) ? M[i * DLD + k] : 0.0;
  }
  if (tid < NBK) lg[tid] = log(pivs[tid] * invs[tid]);
  __syncthreads();
  // ---- inverse, in place: diagonal 16x16 pieces first, then recursive doubling ----
  for (int e = tid; e < NBK * NBK; e += DIAG_THREADS) {
    const int i = e >> 7, k = e & 127;
    if ((i >> 4) == (k >> 4)) M[i * DLD + k] = DS[i * DSLD + (k & 15)];       // includes the zero upper part
    else if (k > i) M[i * DLD + k] = 0.0;
  }
  __syncthreads();
  for (int h = SB; h < NBK; h <<= 1) {
    const int fpr = h / 8, n_pairs = NBK / (2 * h), per_pair = fpr * fpr;
    // T = B * A^-1   (A^-1 lower triangular: k >= column fragment start)
    for (int f = warp; f < n_pairs * per_pair; f += DIAG_WARPS) {
      const int t = f / per_pair, fi = (f % per_pair) / fpr, fn = f % fpr, base = 2 * h * t;
      double c0 = 0.0, c1 = 0.0;
      frag_mac<true>(c0, c1, M + (base + h + 8 * fi) * DLD + base, DLD, M + base * DLD + base + 8 * fn, DLD, 8 * fn, h, 1.0, lane);
      *reinterpret_cast<double2*>(T + (t * h + 8 * fi + r) * TLD + 8 * fn + 2 * c) = make_double2(c0, c1);
    }
    __syncthreads();
    // X21 = -C^-1 * T  (C^-1 lower triangular: k < row fragment end)
    for (int f = warp; f < n_pairs * per_pair; f += DIAG_WARPS) {
      const int t = f / per_pair, fi = (f % per_pair) / fpr, fn = f % fpr, base = 2 * h * t;
      double c0 = 0.0, c1 = 0.0;
      frag_mac<true>(c0, c1, M + (base + h + 8 * fi) * DLD + base + h, DLD, T + (t * h) * TLD + 8 * fn, TLD, 0, 8 * fi + 8, -1.0, lane);
      *reinterpret_cast<double2*>(M + (base + h + 8 * fi + r) * DLD + base + 8 * fn + 2 * c) = make_double2(c0, c1);
    }
    __syncthreads();
  }
  // x_j = inv(L_jj) r_j ; outputs
  if (tid < NBK) {
    double xv = 0.0;
    for (int k = 0; k <= tid; ++k) xv = fma(M[tid * DLD + k], rj[k], xv);
    beta[(long long)s * Np + j * NBK + tid] = xv;
  }
  double* di = dinv + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  double* dt = dinvT + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  for (int e = tid; e < NBK * NBK; e += DIAG_THREADS) {   // both in tiled order, ready to be GEMM operands
    const int kk = e & 3, i = (e >> 2) & 127, k = (e >> 11) * TILE_K + ((e >> 9) & 3) * 4 + kk;
    di[e] = (k <= i) ? M[i * DLD + k] : 0.0;
    dt[e] = (k >= i) ? M[k * DLD + i] : 0.0;
  }
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
  extern __shared__ __align__(128) double smem[];
  const int s = s0 + blockIdx.z, it = j + 1 + blockIdx.x, KPB = NBK / TILE_K;
  double* tile = fac + (long long)s * fac_stride + tile_off(Np / TILE_K, it, j * KPB);
  const double* B = dinv + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  Ring ring; ring.init(smem);
  Acc acc; acc.zero();
  mainloop_bulk(ring, tile, B, KPB, acc);
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
      *reinterpret_cast<double2*>(tile + elem_off(row, col)) = make_double2(acc.c[i][jj][0], acc.c[i][jj][1]);
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
  extern __shared__ __align__(128) double smem[];
  const int it = it0 + blockIdx.x, nt = nt0 + blockIdx.y;
  if (nt > it) return;
  const int s = s0 + blockIdx.z, KPB = NBK / TILE_K, KTA = Np / TILE_K;
  double* base = fac + (long long)s * fac_stride;
  const double* A = base + tile_off(KTA, it, kb0 * KPB);
  const double* B = base + tile_off(KTA, nt, kb0 * KPB);
  double* C = base + tile_off(KTA, it, nt * KPB);
  Ring ring; ring.init(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
  if (threadIdx.x == 0)   // start the operand stream before touching C
    for (int p = 0; p < 2 && p < (kb1 - kb0) * KPB; ++p) ring.produce(A + (long long)p * TILE_DOUBLES, B + (long long)p * TILE_DOUBLES);
  // acc starts at -C (loaded while the first bulk copies are in flight), accumulates +A B^T, and C_new = -acc:
  // the read of C is off the critical path and the epilogue is stores only
  Acc acc;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const double2 v = *reinterpret_cast<const double2*>(C + elem_off(frag_row(wm, i, lane), frag_col(wn, jj, lane)));
      acc.c[i][jj][0] = -v.x;
      acc.c[i][jj][1] = -v.y;
    }
  mainloop_bulk(ring, A, B, (kb1 - kb0) * KPB, acc, 2);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      *reinterpret_cast<double2*>(C + elem_off(frag_row(wm, i, lane), frag_col(wn, jj, lane))) =
          make_double2(-acc.c[i][jj][0], -acc.c[i][jj][1]);
}

// ---- inversion sweep ------------------------------------------------------------------------------
__global__ void place_diag_kernel(double* __restrict__ fac, long long fac_stride, int Np, const double* __restrict__ dinv,
                                  long long dinv_stride, int s0) {
  const int s = s0 + blockIdx.y, j = blockIdx.x;
  double* blk = fac + (long long)s * fac_stride + tile_off(Np / TILE_K, j, j * (NBK / TILE_K));
  const double* di = dinv + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  for (int e = threadIdx.x; e < NBK * NBK; e += blockDim.x) blk[e] = di[e];   // same tiled order on both sides
}

// T = L[it, j] * inv(L_jj), written transposed into the one-row-block tiled matrix tt (128 x Np):
// tt(c, it*128 + row) = T(row, c)
__global__ void __launch_bounds__(THREADS, 1)
inv_step1_kernel(const double* __restrict__ fac, long long fac_stride, int Np, int j, const double* __restrict__ dinvT,
                 long long dinv_stride, double* __restrict__ tt, long long tt_stride, int s0) {
  extern __shared__ __align__(128) double smem[];
  const int s = s0 + blockIdx.z, it = j + 1 + blockIdx.x, KPB = NBK / TILE_K;
  const double* A = fac + (long long)s * fac_stride + tile_off(Np / TILE_K, it, j * KPB);
  const double* B = dinvT + (long long)s * dinv_stride + (long long)j * NBK * NBK;
  Ring ring; ring.init(smem);
  Acc acc; acc.zero();
  mainloop_bulk(ring, A, B, KPB, acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
  double* T = tt + (long long)s * tt_stride;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int k = it * NBK + frag_row(wm, i, lane), col = frag_col(wn, jj, lane);
      T[elem_off(col, k)] = acc.c[i][jj][0];
      T[elem_off(col + 1, k)] = acc.c[i][jj][1];
    }
}

// X[it, j] = - sum_{k=(j+1)*128}^{(it+1)*128-1} Linv[it, k] * T[k, :]
__global__ void __launch_bounds__(THREADS, 1)
inv_step2_kernel(double* __restrict__ fac, long long fac_stride, int Np, int j, const double* __restrict__ tt,
                 long long tt_stride, int s0) {
  extern __shared__ __align__(128) double smem[];
  // longest rows first: better tail behaviour
  const int n_it = gridDim.x, it = j + n_it - (int)blockIdx.x, KPB = NBK / TILE_K, KTA = Np / TILE_K;
  const int s = s0 + blockIdx.z;
  double* base = fac + (long long)s * fac_stride;
  const double* A = base + tile_off(KTA, it, (j + 1) * KPB);
  const double* B = tt + (long long)s * tt_stride + (long long)(j + 1) * KPB * TILE_DOUBLES;
  double* C = base + tile_off(KTA, it, j * KPB);
  Ring ring; ring.init(smem);
  Acc acc; acc.zero();
  mainloop_bulk(ring, A, B, (it - j) * KPB, acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      *reinterpret_cast<double2*>(C + elem_off(frag_row(wm, i, lane), frag_col(wn, jj, lane))) =
          make_double2(-acc.c[i][jj][0], -acc.c[i][jj][1]);
}

// tiled <-> row-major conversion of the N x N leading part (tests, b7_gp_read_factor)
__global__ void untile_kernel(const double* __restrict__ facT, double* __restrict__ out, int Np, int N) {
  const int row = blockIdx.x;
  const double* src = facT + tile_off(Np / TILE_K, row >> 7, 0);
  for (int k = threadIdx.x; k < N; k += blockDim.x) out[(long long)row * N + k] = src[elem_off(row & 127, k)];
}

bool g_attr_done[16] = {false};   // function attributes are per device
int set_attrs(int device) {
  if (g_attr_done[device & 15]) return 0;
  B7_CUDA(cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
  B7_CUDA(cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RING_SMEM));
  B7_CUDA(cudaFuncSetAttribute(trail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RING_SMEM));
  B7_CUDA(cudaFuncSetAttribute(inv_step1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RING_SMEM));
  B7_CUDA(cudaFuncSetAttribute(inv_step2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RING_SMEM));
  g_attr_done[device & 15] = true;
  return 0;
}

// ---- alpha = L^-T beta ------------------------------------------------------------------------------
// The INT8 posterior path takes the mean from an fp64 dot product k*^T alpha (the arithmetic the oracle uses,
// `Ks @ alpha`) instead of v^T beta through the sliced operands, so alpha is needed once per fit.  Blocked back
// substitution on the factor before it is inverted in place: column sweep from the bottom, one CTA per draw,
//   alpha_ib = inv(L_ib,ib)^T r_ib ;   r[k] -= sum_row L[ib*128 + row][k] alpha_ib[row]   for k < 128 ib.
// Row block ib of L is one contiguous run of tiles; a warp owns a k-group (128 rows x 4 columns = 4 KB), lanes are
// rows, and the 4 column sums come out of a fixed shuffle tree (deterministic).
// FROM_INVERSE: fac already holds L^-1 (the path was switched after the fit, or the factor came from another
// rank): alpha = (L^-1)^T beta, the same row-block primitive without the solve.
constexpr int ALPHA_THREADS = 1024;

template <bool FROM_INVERSE>
__global__ void __launch_bounds__(ALPHA_THREADS, 1)
alpha_kernel(const double* __restrict__ fac, long long fac_stride, int Np, const double* __restrict__ dinvT, long long dinv_stride,
             const double* __restrict__ beta, double* __restrict__ alpha, int s0) {
  extern __shared__ __align__(16) double sm[];
  double* r = sm;            // Np
  double* xa = sm + Np;      // 128
  const int s = s0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NB = Np / NBK;
  const double* F = fac + (long long)s * fac_stride;
  for (int e = tid; e < Np; e += ALPHA_THREADS) r[e] = FROM_INVERSE ? 0.0 : beta[(long long)s * Np + e];
  __syncthreads();
  for (int step = 0; step < NB; ++step) {
    const int ib = FROM_INVERSE ? step : NB - 1 - step;
    if (tid < NBK) {
      if (FROM_INVERSE) {
        xa[tid] = beta[(long long)s * Np + ib * NBK + tid];
      } else {
        const double* dt = dinvT + (long long)s * dinv_stride + (long long)ib * NBK * NBK;
        double sum = 0.0;
#pragma unroll 4
        for (int q4 = 0; q4 < NBK / 4; ++q4) {
          const double4 m = *reinterpret_cast<const double4*>(dt + (q4 >> 2) * TILE_DOUBLES + (q4 & 3) * (BM * 4) + tid * 4);
          const double* rr = r + ib * NBK + 4 * q4;
          sum = fma(m.x, rr[0], sum); sum = fma(m.y, rr[1], sum); sum = fma(m.z, rr[2], sum); sum = fma(m.w, rr[3], sum);
        }
        xa[tid] = sum;
        alpha[(long long)s * Np + ib * NBK + tid] = sum;
      }
    }
    __syncthreads();
    const double* rowblk = F + tile_off(Np / TILE_K, ib, 0);
    const int n_kg = (FROM_INVERSE ? ib + 1 : ib) * (NBK / 4);       // k-groups of 4 columns left of (or including) the diagonal block
    for (int kg = warp; kg < n_kg; kg += ALPHA_THREADS / 32) {
      const double* p = rowblk + (long long)kg * (BM * 4);
      double4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const double4*>(p + (lane + 32 * i) * 4);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double x = xa[lane + 32 * i];
        a0 = fma(v[i].x, x, a0); a1 = fma(v[i].y, x, a1); a2 = fma(v[i].z, x, a2); a3 = fma(v[i].w, x, a3);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
      }
      if (lane == 0) {
        double* out = r + kg * 4;
        if (FROM_INVERSE) { out[0] += a0; out[1] += a1; out[2] += a2; out[3] += a3; }
        else { out[0] -= a0; out[1] -= a1; out[2] -= a2; out[3] -= a3; }
      }
    }
    __syncthreads();
  }
  if (FROM_INVERSE)
    for (int e = tid; e < Np; e += ALPHA_THREADS) alpha[(long long)s * Np + e] = r[e];
}

}  // namespace

int b7_launch_potrf(b7_gp* gp, int s0, int count) {
  b7_ctx* ctx = gp->ctx;
  B7_CHECK(set_attrs(ctx->device));
  const int Np = gp->Np, NB = gp->NB;
  const long long fs = (long long)Np * Np, ds = (long long)NB * NBK * NBK;
  static const int W_env = getenv("B7_POTRF_W") ? atoi(getenv("B7_POTRF_W")) : 0;
  const int W = W_env > 0 ? W_env : 4;   // outer panel = 4 blocks (512 columns)
  cudaStream_t sa = ctx->stream, sb = ctx->stream2;
  bool far_pending = false;
  // INT8 path (potrf_i8.cu): the panel rows are sliced once per outer panel into one of two scratch sets (the far
  // update of panel J still reads its set while panel J + W is sliced)
  // (batched fits only: a single factor is a latency chain that the extra slicing launch makes longer; measured
  // at N = 4096: S = 32 26.7 -> 17.7 ms, S = 1 4.1 -> 4.8 ms)
  // refit latency, ms, int8 / fp64 updates: N = 4096 S = 4 5.7 / 5.9, S = 8 7.4 / 8.8, S = 16 11.2 / 15.3;
  // N = 2048 S = 4 2.5 / 2.25, S = 16 3.5 / 3.5, S = 32 4.3 / 4.8 -> needs count * NB^2 >= 4096
  // The choice looks at the handle's total number of draws, not at how many of them this call factorises: a sharded fit
  // (each GPU factorises S / G draws) then uses the same arithmetic as the one-GPU fit and stays bit-identical to it.
  const int batch = gp->S;
  const bool i8 = ctx->use_i8 && ctx->potrf_i8 && batch >= 4 && (long long)batch * NB * NB >= 4096 && NB > W && Np <= B7_I8_MAX_NP;
  int8_t* pS[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  double* pSig[2] = {nullptr, nullptr};
  const size_t p_stride = i8 ? b7_i8_panel_bytes(Np, W) : 0;
  if (i8)
    for (int b = 0; b < 2; ++b) {
      B7_CHECK(b7_pool_alloc(ctx, (void**)&pS[b][0], p_stride * count));
      B7_CHECK(b7_pool_alloc(ctx, (void**)&pS[b][1], p_stride * count));
      B7_CHECK(b7_pool_alloc(ctx, (void**)&pSig[b], sizeof(double) * (size_t)Np * count));
    }
  int set = 0;
  for (int J = 0; J < NB; J += W) {
    const int Jend = J + W < NB ? J + W : NB;
    // --- the panel's own columns: latency-bound chain on the main stream ---
    for (int j = J; j < Jend; ++j) {
      diag_kernel<<<count, DIAG_THREADS, DIAG_SMEM, sa>>>(gp->fac, fs, Np, j, gp->dinv, gp->dinvT, ds, gp->beta,
                                                        gp->logdet, gp->info, s0);
      b7_count(ctx);
      const int rem = NB - 1 - j;
      if (rem > 0) {
        panel_kernel<<<dim3(rem, 1, count), THREADS, RING_SMEM, sa>>>(gp->fac, fs, Np, j, gp->dinv, ds, gp->beta, s0);
        b7_count(ctx);
      }
      if (Jend - 1 - j > 0) {   // columns j+1 .. Jend-1 of the outer panel, all rows below
        trail_kernel<<<dim3(rem, Jend - 1 - j, count), THREADS, RING_SMEM, sa>>>(gp->fac, fs, Np, j, j + 1, j + 1, j + 1, s0);
        b7_count(ctx);
      }
    }
    if (Jend >= NB) break;
    // --- update by this panel (k = 512).  "near": the next panel's columns, needed at once, main stream
    //     (after the previous far update, which touched the same tiles).  "far": everything to the right of
    //     the next panel, on the second stream, overlapping the next panel's chain of small kernels. ---
    const int near_end = Jend + W < NB ? Jend + W : NB;
    if (i8) B7_CHECK(b7_i8_panel_slice(ctx, sa, gp->fac, Np, J, Jend, Jend, pS[set][0], pS[set][1], p_stride, pSig[set], s0, count));
    if (far_pending) B7_CUDA(cudaStreamWaitEvent(sa, ctx->evB, 0));
    if (i8) {
      B7_CHECK(b7_i8_trail(ctx, sa, gp->fac, Np, pS[set][0], pS[set][1], p_stride, pSig[set], J, Jend, Jend, Jend, NB - Jend, Jend,
                           near_end - Jend, s0, count));
    } else {
      trail_kernel<<<dim3(NB - Jend, near_end - Jend, count), THREADS, RING_SMEM, sa>>>(gp->fac, fs, Np, J, Jend, Jend, Jend, s0);
      b7_count(ctx);
    }
    if (near_end < NB) {
      B7_CUDA(cudaEventRecord(ctx->evA, sa));
      B7_CUDA(cudaStreamWaitEvent(sb, ctx->evA, 0));
      if (i8) {
        B7_CHECK(b7_i8_trail(ctx, sb, gp->fac, Np, pS[set][0], pS[set][1], p_stride, pSig[set], J, Jend, Jend, near_end, NB - near_end,
                             near_end, NB - near_end, s0, count));
      } else {
        trail_kernel<<<dim3(NB - near_end, NB - near_end, count), THREADS, RING_SMEM, sb>>>(gp->fac, fs, Np, J, Jend, near_end, near_end, s0);
        b7_count(ctx);
      }
      B7_CUDA(cudaEventRecord(ctx->evB, sb));
      far_pending = true;
    }
    set ^= 1;
  }
  if (far_pending) B7_CUDA(cudaStreamWaitEvent(sa, ctx->evB, 0));
  if (i8)
    for (int b = 0; b < 2; ++b) {
      b7_pool_free(ctx, pS[b][0]);
      b7_pool_free(ctx, pS[b][1]);
      b7_pool_free(ctx, pSig[b]);
    }
  B7_CUDA(cudaGetLastError());
  return 0;
}

// alpha = L^-T beta for draws [s0, s0 + count): from L (before the inversion) or from L^-1 (from_inverse)
int b7_launch_alpha(b7_gp* gp, int s0, int count, bool from_inverse) {
  b7_ctx* ctx = gp->ctx;
  const int Np = gp->Np;
  const size_t smem = (size_t)(Np + NBK) * sizeof(double);
  if (smem > 200 * 1024) { b7_set_error("alpha: %d observations exceed the shared-memory sweep", Np); return B7_ERR_ARG; }
  static bool attr[16][2] = {{false}};
  if (!attr[ctx->device & 15][from_inverse]) {
    if (from_inverse) B7_CUDA(cudaFuncSetAttribute(alpha_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    else B7_CUDA(cudaFuncSetAttribute(alpha_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr[ctx->device & 15][from_inverse] = true;
  }
  const long long fs = (long long)Np * Np, ds = (long long)gp->NB * NBK * NBK;
  if (from_inverse) alpha_kernel<true><<<count, ALPHA_THREADS, smem, ctx->stream>>>(gp->fac, fs, Np, gp->dinvT, ds, gp->beta, gp->alpha, s0);
  else alpha_kernel<false><<<count, ALPHA_THREADS, smem, ctx->stream>>>(gp->fac, fs, Np, gp->dinvT, ds, gp->beta, gp->alpha, s0);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

int b7_launch_untile(b7_ctx* ctx, const double* facT, double* out, int Np, int N) {
  untile_kernel<<<N, 256, 0, ctx->stream>>>(facT, out, Np, N);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}

int b7_launch_trtri(b7_gp* gp, int s0, int count) {
  b7_ctx* ctx = gp->ctx;
  B7_CHECK(set_attrs(ctx->device));
  const int Np = gp->Np, NB = gp->NB;
  const long long fs = (long long)Np * Np, ds = (long long)NB * NBK * NBK, ts = (long long)NBK * Np;
  place_diag_kernel<<<dim3(NB, count), 256, 0, ctx->stream>>>(gp->fac, fs, Np, gp->dinv, ds, s0);
  b7_count(ctx);
  // INT8 path: block-recursive inversion with static operands per level (trtri_i8.cu)
  if (ctx->use_i8 && ctx->trtri_i8 && Np <= B7_I8_MAX_NP) return b7_launch_trtri_i8(gp, s0, count);
  for (int j = NB - 2; j >= 0; --j) {
    const int rem = NB - 1 - j;
    inv_step1_kernel<<<dim3(rem, 1, count), THREADS, RING_SMEM, ctx->stream>>>(gp->fac, fs, Np, j, gp->dinvT, ds, gp->tt, ts, s0);
    inv_step2_kernel<<<dim3(rem, 1, count), THREADS, RING_SMEM, ctx->stream>>>(gp->fac, fs, Np, j, gp->tt, ts, s0);
    b7_count(ctx, 2);
  }
  B7_CUDA(cudaGetLastError());
  return 0;
}
