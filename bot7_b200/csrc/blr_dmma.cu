// DNGO candidate pass on the FP64 tensor pipe: (optional) ReLU MLP basis followed by the BLR head, tile by tile.
//
// Replaces, for the candidate grid, what models/dngo.lua:155-174 does in 32-row Lua minibatches: Z1 = basis(X_hid)
// (:165-171) and predictor:predict(Z0, Y0, Z1, ...) (:174).  One CTA works on tiles of 128 candidates that never
// leave shared memory between the layers:
//     act_0 = the tile's rows of the input grid (X, or precomputed features)
//     act_l+1 = act(W_l act_l + b_l)                      l = 0 .. n_layers-1   (nn.Linear + ReLU)
//     v = L_A^-1 phi,  var = |v|^2 + 1/beta,  mean = m + w^T phi                 per (alpha_p, beta) draw
// Every product is a DMMA m8n8k4 fragment product: the weights (and, for the head, [L_A^-1 ; w^T] as one matrix whose
// last row yields the mean) sit in shared memory for the whole launch with a leading dimension = 4 mod 16 doubles
// (conflict-free fragment loads), a warp owns 16 candidates and all output rows, and the head skips the fragments
// above the diagonal of the triangular L_A^-1.  The one-DFMA-thread-per-candidate kernels of blr.cu ran at 27 %
// (BLR) and 22 % (MLP) of the FP64 pipe; they remain as the fallback for shapes whose weights do not fit.
// With the basis fused, the 8 D bytes per candidate of Z1 (1.7 GB at config 4) are never written or read.
#include <algorithm>
#include <vector>

#include "b7_internal.h"
#include "gemm_tile.cuh"

using b7g::dmma884;

namespace {

constexpr int TC = 128;            // candidates per tile
constexpr int THREADS = 256;       // 8 warps x 16 candidates
constexpr int MAXL = 4;            // MLP layers
constexpr int MAXF = 8;            // output fragments of 8 rows (widths <= 64)

// leading dimension: >= width rounded up to 4, = 4 mod 16 (fragment rows land in distinct banks)
__host__ __device__ inline int ld_of(int width) {
  const int w4 = (width + 3) / 4 * 4;
  return w4 + ((4 - w4 % 16) + 16) % 16;
}

struct Plan {
  int n_layers;                 // MLP layers before the head
  int width[MAXL + 1];          // width[0] = input width, width[l + 1] = output of layer l
  int relu[MAXL];
  int w_off[MAXL];              // offsets (doubles) of the staged weights / biases in shared memory
  int b_off[MAXL];
  int head_off;                 // [S][rows_pad][ld(D)] head matrices
  int head_rows;                // D + 1 rounded up to 8
  int act_off;                  // activation buffer [TC][act_ld]; a layer overwrites it in place (a warp owns its 16 rows)
  int act_ld;
  int S;                        // draws staged (0: no head, write the features)
};

struct Ptrs {
  const double* W[MAXL];
  const double* b[MAXL];
};

// acc[i][j] (+)= A[8i.., k] * B[cand.., k]^T over k < K4; A: rows x lda, Bt: [cand][ldb].
// MF (output fragments) is a compile-time constant so that the fragment loop carries no branches: the loads of a pair of
// k-steps (up to 2 x 7 A fragments + 4 B fragments) are issued together in front of their up to 28 DMMAs (with a runtime
// bound and `break` in the loop every fragment load sat in front of its own two DMMAs: 28 % of the DMMA issue rate).
// tri: A is [L^-1 ; w^T] -- fragment i is zero for k > 8 i + 7, except the last one (it holds the dense row w^T), so the
// fragments active at k0 are the suffix i >= k0 / 8.
template <int MF>
__device__ __forceinline__ void tile_mma_t(const double* __restrict__ A, int lda, int K4, const double* __restrict__ Bt, int ldb, int lane,
                                           double (&acc)[MAXF][2][2], bool tri) {
  const int r = lane >> 2, c = lane & 3;
  const double* Ar = A + r * lda + c;
  const double* B0 = Bt + r * ldb + c;
  const double* B1 = Bt + (8 + r) * ldb + c;
  // Fragments i >= first are active at this pair of k-steps.  The skip is a real (warp-uniform) jump into a fall-through
  // switch: a predicated-off DMMA still holds the FP64 tensor pipe for its 16 cycles (ncu dngo_r02: 47.7 M DMMAs issued,
  // 28.8 M predicated on, pipe busy for all of them).
#define B7_FRAG_LOAD(I, KK)                                                  \
  if (MF > I) { av[I] = Ar[8 * I * lda + KK]; }
#define B7_FRAG_MMA(I, BA, BB)                                               \
  if (MF > I) { dmma884(acc[I][0][0], acc[I][0][1], av[I], BA); dmma884(acc[I][1][0], acc[I][1][1], av[I], BB); }
#define B7_FRAG_SWITCH(OP, ...)                                              \
  switch (first) {                                                           \
    case 0: OP(0, ##__VA_ARGS__) case 1: OP(1, ##__VA_ARGS__) case 2: OP(2, ##__VA_ARGS__) case 3: OP(3, ##__VA_ARGS__)   \
    case 4: OP(4, ##__VA_ARGS__) case 5: OP(5, ##__VA_ARGS__) case 6: OP(6, ##__VA_ARGS__) default: OP(7, ##__VA_ARGS__)  \
  }
  // one k-step (4 columns) at a time: the A fragments of the step, then its DMMAs (14 independent accumulators between two
  // DMMAs on the same one); 16 warps per SM (two CTAs) cover the load-to-use latency
  for (int k0 = 0; k0 < K4; k0 += 4) {
    const int imin = tri ? k0 >> 3 : 0, first = imin < MF - 1 ? imin : MF - 1;
    const double b0 = B0[k0], b1 = B1[k0];
    double av[8];
    B7_FRAG_SWITCH(B7_FRAG_LOAD, k0)
    B7_FRAG_SWITCH(B7_FRAG_MMA, b0, b1)
  }
#undef B7_FRAG_LOAD
#undef B7_FRAG_MMA
#undef B7_FRAG_SWITCH
}

__device__ __forceinline__ void tile_mma(const double* __restrict__ A, int lda, int mf, int K4, const double* __restrict__ Bt, int ldb,
                                         int lane, double (&acc)[MAXF][2][2], bool tri) {
  switch (mf) {
    case 1: tile_mma_t<1>(A, lda, K4, Bt, ldb, lane, acc, tri); break;
    case 2: tile_mma_t<2>(A, lda, K4, Bt, ldb, lane, acc, tri); break;
    case 3: tile_mma_t<3>(A, lda, K4, Bt, ldb, lane, acc, tri); break;
    case 4: tile_mma_t<4>(A, lda, K4, Bt, ldb, lane, acc, tri); break;
    case 5: tile_mma_t<5>(A, lda, K4, Bt, ldb, lane, acc, tri); break;
    case 6: tile_mma_t<6>(A, lda, K4, Bt, ldb, lane, acc, tri); break;
    case 7: tile_mma_t<7>(A, lda, K4, Bt, ldb, lane, acc, tri); break;
    default: tile_mma_t<8>(A, lda, K4, Bt, ldb, lane, acc, tri); break;
  }
}

__global__ void __launch_bounds__(THREADS, 2)
dngo_tile_kernel(const double* __restrict__ in, long long M, Plan pl, Ptrs pt, const double* __restrict__ Linv, const double* __restrict__ wv,
                 const double* __restrict__ par, int D, long long ld_out, double* __restrict__ mean, double* __restrict__ var,
                 double* __restrict__ feat_out) {
  extern __shared__ __align__(16) double sh[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, c = lane & 3;
  // ---- stage the weights once per CTA (zero padded) ----
  for (int l = 0; l < pl.n_layers; ++l) {
    const int hi = pl.width[l], ho = pl.width[l + 1], ldw = ld_of(hi), rows = (ho + 7) / 8 * 8;
    for (int e = tid; e < rows * ldw; e += THREADS) {
      const int j = e / ldw, k = e % ldw;
      sh[pl.w_off[l] + e] = (j < ho && k < hi) ? pt.W[l][(long long)j * hi + k] : 0.0;
    }
    for (int e = tid; e < rows; e += THREADS) sh[pl.b_off[l] + e] = e < ho ? pt.b[l][e] : 0.0;
  }
  const int ldh = ld_of(D);
  for (int s = 0; s < pl.S; ++s)
    for (int e = tid; e < pl.head_rows * ldh; e += THREADS) {
      const int i = e / ldh, k = e % ldh;
      double v = 0.0;
      if (k < D) {
        if (i < D) v = k <= i ? Linv[((long long)s * D + i) * D + k] : 0.0;
        else if (i == D) v = wv[(long long)s * D + k];
      }
      sh[pl.head_off + s * pl.head_rows * ldh + e] = v;
    }
  __syncthreads();

  const long long n_tiles = (M + TC - 1) / TC;
  const int w_in = pl.width[0], lda = pl.act_ld, w4 = (w_in + 3) / 4 * 4, padw = w4 - w_in;
  double* cur = sh + pl.act_off;
  const int dq = THREADS / w_in, dr = THREADS % w_in, cc0 = tid / w_in, kk0 = tid % w_in;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long c0 = tile * TC;
    const int nc = (int)min((long long)TC, M - c0);
    // ---- input tile: nc x w_in contiguous doubles -> [cand][lda].  All of a thread's global loads are issued before the
    // first shared-memory store (a loop of load / store pairs was latency bound, ~5 us per tile); the second CTA of the SM
    // computes meanwhile.  Zero padding: columns w_in .. w4-1 (a wider layer output of the previous tile may have used
    // them) and rows >= nc ----
    {
      constexpr int NLD = 16;                          // loads in flight per thread; two rounds cover widths <= 64
      const double* src = in + c0 * w_in;
      const int total = nc * w_in;
      int cc = cc0, k = kk0;
#pragma unroll 1
      for (int round = 0; round < TC * 64 / THREADS / NLD; ++round) {
        const int f0 = tid + round * NLD * THREADS;
        if (round * NLD * THREADS >= TC * w_in) break;
        double buf[NLD];
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
          const int f = f0 + i * THREADS;
          buf[i] = f < total ? src[f] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
          if (f0 + i * THREADS < TC * w_in) cur[cc * lda + k] = buf[i];     // rows >= nc receive the zeros loaded above
          cc += dq; k += dr;
          if (k >= w_in) { k -= w_in; ++cc; }
        }
      }
      for (int e = tid; e < TC * padw; e += THREADS) cur[(e / padw) * lda + w_in + e % padw] = 0.0;
    }
    __syncthreads();
    // ---- MLP layers ----
    for (int l = 0; l < pl.n_layers; ++l) {
      const int hi = pl.width[l], ho = pl.width[l + 1], mf = (ho + 7) / 8, K4 = (hi + 3) / 4 * 4;
      double acc[MAXF][2][2];
#pragma unroll
      for (int i = 0; i < MAXF; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
      tile_mma(sh + pl.w_off[l], ld_of(hi), mf, K4, cur + warp * 16 * lda, lda, lane, acc, false);
      const double* bias = sh + pl.b_off[l];
      const int ho4 = (ho + 3) / 4 * 4;
      __syncwarp();                                    // every lane has read the warp's 16 input rows: overwrite them in place
#pragma unroll
      for (int i = 0; i < MAXF; ++i) {
        if (i >= mf) break;
        const int row = 8 * i + r;
        if (row >= ho4) continue;                      // columns ho .. ho4-1 must be written (zeros) for the next K loop
        const double bb = bias[row];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double v = row < ho ? acc[i][j][e] + bb : 0.0;
            if (pl.relu[l]) v = v > 0.0 ? v : 0.0;
            cur[(warp * 16 + 8 * j + 2 * c + e) * lda + row] = v;
          }
      }
      __syncwarp();                                    // a warp only reads the 16 candidates it wrote
    }
    const int w_last = pl.width[pl.n_layers];
    if (pl.S == 0) {
      // ---- features out (b7_mlp_features): [cand][w_last] contiguous ----
      __syncthreads();
      for (int e = tid; e < nc * w_last; e += THREADS) feat_out[c0 * w_last + e] = cur[(e / w_last) * lda + e % w_last];
    } else {
      // ---- BLR head per draw: rows 0 .. D-1 of the product are v = L_A^-1 phi, row D is w^T phi ----
      const int mf = pl.head_rows / 8, K4 = (D + 3) / 4 * 4;      // the dense row w^T sits in the last fragment (D / 8 = mf - 1)
      for (int s = 0; s < pl.S; ++s) {
        double acc[MAXF][2][2];
#pragma unroll
        for (int i = 0; i < MAXF; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
        tile_mma(sh + pl.head_off + s * pl.head_rows * ldh, ldh, mf, K4, cur + warp * 16 * lda, lda, lane, acc, true);
        double s2[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, mu[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int i = 0; i < MAXF; ++i) {
          if (i >= mf) break;
          const int row = 8 * i + r;
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const double v = acc[i][j][e];
              if (row < D) s2[j][e] = fma(v, v, s2[j][e]);
              else if (row == D) mu[j][e] = v;
            }
        }
        // rows are spread over the 8 lanes that share lane & 3: fixed xor tree
#pragma unroll
        for (int o = 4; o < 32; o <<= 1)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              s2[j][e] += __shfl_xor_sync(0xffffffffu, s2[j][e], o);
              mu[j][e] += __shfl_xor_sync(0xffffffffu, mu[j][e], o);
            }
        if (r == 0) {
          const double mconst = par[s * 4 + 2], inv_beta = par[s * 4 + 3];
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = warp * 16 + 8 * j + 2 * c + e;
              if (cc < nc) {
                mean[(long long)s * ld_out + c0 + cc] = mconst + mu[j][e];
                var[(long long)s * ld_out + c0 + cc] = s2[j][e] + inv_beta;
              }
            }
        }
      }
    }
    __syncthreads();                                   // the activation buffers are refilled by the next tile
  }
}

}  // namespace

// Plans the shared memory, uploads nothing (all pointers are device pointers) and launches.  Returns 1 if the shapes
// do not fit the tile kernel (the caller then uses the per-candidate kernels of blr.cu).
int b7_launch_dngo_tiles(b7_ctx* ctx, const double* in, int64_t M, int n_layers, const int* dims, const double* const* W_dev,
                         const double* const* b_dev, int relu_last, const double* Linv, const double* w, const double* par, int D, int S,
                         int64_t ld_out, double* mean, double* var, double* feat_out) {
  if (M <= 0) return 0;
  if (n_layers > MAXL) return 1;
  Plan pl = {};
  Ptrs pt = {};
  pl.n_layers = n_layers;
  pl.S = S;
  int off = 0, wmax = dims[0];
  for (int l = 0; l <= n_layers; ++l) {
    if (dims[l] < 1 || dims[l] > 64) return 1;
    pl.width[l] = dims[l];
    wmax = std::max(wmax, dims[l]);
  }
  if (S > 0 && (D != dims[n_layers] || D + 1 > 64)) return 1;
  for (int l = 0; l < n_layers; ++l) {
    const int rows = (dims[l + 1] + 7) / 8 * 8;
    pl.relu[l] = (l < n_layers - 1 || relu_last) ? 1 : 0;
    pl.w_off[l] = off; off += rows * ld_of(dims[l]);
    pl.b_off[l] = off; off += rows;
    pt.W[l] = W_dev[l];
    pt.b[l] = b_dev[l];
  }
  pl.head_rows = S > 0 ? (D + 1 + 7) / 8 * 8 : 0;
  pl.head_off = off; off += S * pl.head_rows * (S > 0 ? ld_of(D) : 0);
  pl.act_ld = ld_of(wmax);
  pl.act_off = off; off += TC * pl.act_ld;
  const size_t smem = (size_t)off * sizeof(double);
  if (smem > 226 * 1024) return 1;
  static bool done[16] = {false};
  if (!done[ctx->device & 15]) {
    B7_CUDA(cudaFuncSetAttribute(dngo_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    done[ctx->device & 15] = true;
  }
  const int64_t n_tiles = (M + TC - 1) / TC;
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * per_sm);
  dngo_tile_kernel<<<grid, THREADS, smem, ctx->stream>>>(in, M, pl, pt, Linv, w, par, D, ld_out, mean, var, feat_out);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}
