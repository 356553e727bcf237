// Sobol candidate grid from integer direction numbers (kernel #4).
//
// Replaces grid:i4_sobol / grid:generate (reference grids/sobol.lua:58-90,216-335), where every
// coordinate costs one emulated XOR built from ~100 tiny tensor ops (utils/bits.lua:79-82).
// The state machine of i4_sobol reduces to the closed form
//     lastq_i(seed) = XOR_{b in bits(gray(seed))} V[i][b],   gray(s) = s ^ (s >> 1),
// so every point is independent.  Work is split into aligned tiles of 4096 consecutive seeds:
// the contribution of gray bits >= 12 is constant over a tile (one XOR chain per dimension per
// tile), the low 12 bits go through three 16-entry tables per dimension.  The output is written
// straight in row-major order, one fp64 per lane: the kernel is bound by the HBM write,
// 8 B per coordinate, nothing is read.
#include "b7_internal.h"

namespace {

__constant__ uint32_t c_dirs[B7_MAX_DIMS * B7_SOBOL_BITS];

// primitive polynomials (Bratley & Fox), as listed in grids/sobol.lua:46-52
const int kPoly[B7_MAX_DIMS] = {1,   3,   7,   11,  13,  19,  25,  37,  59,  47,  61,  55,  41,  67,
                                97,  91,  109, 103, 115, 131, 193, 137, 145, 143, 241, 157, 185, 167,
                                229, 171, 213, 191, 253, 203, 211, 239, 247, 285, 369, 299};
// initial direction numbers m_{i,j}: table[j][...] holds column j+2 (1-based), for rows
// first_row[j].. (1-based), as listed in grids/sobol.lua:344-390
const int kCol2[] = {1, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3, 1,
                     3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3, 1, 3, 1, 3};
const int kCol3[] = {7, 5, 1, 3, 3, 7, 5, 5, 7, 7, 1, 3, 3, 7, 5, 1, 1, 5, 3,
                     3, 1, 7, 5, 1, 3, 3, 7, 5, 1, 1, 5, 7, 7, 5, 1, 3, 3};
const int kCol4[] = {1, 7, 9,  13, 11, 1, 3,  7, 9,  5,  13, 13, 11, 3, 15, 5, 3, 15,
                     7, 9, 13, 9,  1,  11, 7, 5, 15, 1,  15, 11, 5,  3, 1,  7, 9};
const int kCol5[] = {9,  3,  27, 15, 29, 21, 23, 19, 11, 25, 7,  13, 17, 1, 25, 29, 3,
                     31, 11, 5,  23, 27, 19, 21, 5,  1,  17, 13, 7,  15, 9, 31, 9};
const int kCol6[] = {37, 33, 7, 5,  11, 39, 63, 27, 17, 15, 23, 29, 3, 21,
                     13, 31, 25, 9, 49, 33, 19, 29, 11, 19, 27, 15, 25};
const int kCol7[] = {13, 33, 115, 41, 79, 17, 29, 119, 75, 73, 105, 7, 59, 65, 21, 3, 113, 61, 89, 45, 107};
const int kCol8[] = {7, 23, 39};

struct ColInit { const int* v; int n; int first_row; };
const ColInit kInit[] = {{kCol2, 38, 3}, {kCol3, 37, 4}, {kCol4, 35, 6}, {kCol5, 33, 8},
                         {kCol6, 27, 14}, {kCol7, 21, 20}, {kCol8, 3, 38}};

constexpr int kLowBits = 12;
constexpr int kTile = 1 << kLowBits;

__global__ void __launch_bounds__(256)
sobol_kernel(double* __restrict__ out, int dims, long long first_seed, long long count,
             const double* __restrict__ mins, const double* __restrict__ scale, int affine,
             long long first_tile, long long n_tiles) {
  __shared__ uint32_t tab[3][16][B7_MAX_DIMS];
  __shared__ uint32_t hi_part[B7_MAX_DIMS];
  __shared__ double s_min[B7_MAX_DIMS], s_scale[B7_MAX_DIMS];
  const int tid = threadIdx.x;
  for (int e = tid; e < 3 * 16 * dims; e += blockDim.x) {
    int i = e % dims, r = e / dims, nib = r & 15, t = r >> 4;
    uint32_t x = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (nib & (1 << b)) x ^= c_dirs[i * B7_SOBOL_BITS + t * 4 + b];
    tab[t][nib][i] = x;
  }
  if (tid < dims) {
    s_min[tid] = affine ? mins[tid] : 0.0;
    s_scale[tid] = affine ? scale[tid] : 1.0;
  }
  const float inv_d = 1.0f / (float)dims;
  const long long last_seed = first_seed + count;   // exclusive
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long hi = first_tile + tile;          // seed >> 12
    __syncthreads();
    if (tid < dims) {
      uint32_t g = (uint32_t)(hi ^ (hi >> 1));       // gray bits >= 12 of every seed in the tile
      uint32_t x = 0;
      while (g) {
        int b = __ffs(g) - 1;
        g &= g - 1;
        x ^= c_dirs[tid * B7_SOBOL_BITS + kLowBits + b];
      }
      hi_part[tid] = x;
    }
    __syncthreads();
    const uint32_t carry = ((uint32_t)hi & 1u) << (kLowBits - 1);   // bit 11 of gray also sees bit 12 of seed
    long long s_begin = hi << kLowBits, s_end = s_begin + kTile;
    int lo_begin = (int)((first_seed > s_begin ? first_seed : s_begin) - s_begin);
    int lo_end = (int)((last_seed < s_end ? last_seed : s_end) - s_begin);
    const int n_elem = (lo_end - lo_begin) * dims;
    double* o = out + (s_begin + lo_begin - first_seed) * (long long)dims;
    for (int e = tid; e < n_elem; e += blockDim.x) {
      int p = __float2int_rd(((float)e + 0.5f) * inv_d);   // exact: e < 2^18
      int i = e - p * dims;
      uint32_t lo = (uint32_t)(lo_begin + p);
      uint32_t g = (lo ^ (lo >> 1)) ^ carry;
      uint32_t x = hi_part[i] ^ tab[0][g & 15][i] ^ tab[1][(g >> 4) & 15][i] ^ tab[2][(g >> 8) & 15][i];
      double q = __uint2double_rn(x) * 9.31322574615478515625e-10;   // * 2^-30, exact (sobol.lua:287,329)
      if (affine) q = __dadd_rn(__dmul_rn(q, s_scale[i]), s_min[i]);  // two rounded ops (sobol.lua:79-81)
      o[e] = q;
    }
  }
}

}  // namespace

void b7_sobol_directions_host(int dims, uint32_t* out) {
  // Direction numbers: recurrence of Bratley & Fox section 2 as driven by grids/sobol.lua:236-288.
  long long m[B7_MAX_DIMS][B7_SOBOL_BITS] = {};
  for (int i = 0; i < B7_MAX_DIMS; ++i) m[i][0] = 1;
  for (const ColInit& c : kInit)
    for (int t = 0; t < c.n; ++t) m[c.first_row - 1 + t][(int)(&c - kInit) + 1] = c.v[t];
  for (int j = 0; j < B7_SOBOL_BITS; ++j) m[0][j] = 1;
  for (int i = 0; i < dims; ++i) {
    int deg = 0;
    for (int p = kPoly[i] >> 1; p > 0; p >>= 1) ++deg;
    for (int j = deg; j < B7_SOBOL_BITS; ++j) {
      long long v = m[i][j - deg];
      for (int k = 1; k <= deg; ++k)
        if ((kPoly[i] >> (deg - k)) & 1) v ^= (1LL << k) * m[i][j - k];
      m[i][j] = v;
    }
    for (int j = 0; j < B7_SOBOL_BITS; ++j) out[i * B7_SOBOL_BITS + j] = (uint32_t)(m[i][j] << (B7_SOBOL_BITS - 1 - j));
  }
}

int b7_launch_sobol(b7_ctx* ctx, int dims, int64_t first_seed, int64_t count, const double* mins_dev,
                    const double* scale_dev, double* out_dev) {
  static int cached_dims[16] = {0};   // per device: dims rows currently in constant memory
  if (cached_dims[ctx->device & 15] < dims) {
    uint32_t host[B7_MAX_DIMS * B7_SOBOL_BITS] = {0};
    b7_sobol_directions_host(B7_MAX_DIMS - 1, host);
    B7_CUDA(cudaMemcpyToSymbolAsync(c_dirs, host, sizeof(host), 0, cudaMemcpyHostToDevice, ctx->stream));
    cached_dims[ctx->device & 15] = B7_MAX_DIMS - 1;
  }
  if (count <= 0) return 0;
  long long first_tile = first_seed >> kLowBits;
  long long last_tile = (first_seed + count - 1) >> kLowBits;
  long long n_tiles = last_tile - first_tile + 1;
  // balanced: every block gets the same number of tiles (a 150 us kernel cannot afford a ragged last wave)
  long long want = (long long)ctx->sm_count * 8;
  long long per = (n_tiles + want - 1) / want;
  int grid = (int)((n_tiles + per - 1) / per);
  sobol_kernel<<<grid, 256, 0, ctx->stream>>>(out_dev, dims, first_seed, count, mins_dev, scale_dev,
                                              mins_dev != nullptr, first_tile, n_tiles);
  b7_count(ctx);
  B7_CUDA(cudaGetLastError());
  return 0;
}
