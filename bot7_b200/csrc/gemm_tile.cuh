// FP64 tensor-core tile engine shared by the Cholesky, the triangular inversion and the posterior
// pass:  acc(128x128) += A(128 x K) * B(128 x K)^T, both operands row-major with k contiguous.
//
// sm_100a has no tcgen05 kind for f64 and wgmma does not exist; the FP64 tensor path is warp-level
// mma.sync.m8n8k4.f64, which ptxas lowers to DMMA.8x8x4 (one per 16 clk per SM sub-partition;
// measured 37.1 TFLOP/s on B200, tools/fp64_peak.cu).  A CTA of 8 warps owns a 128x128 tile, a warp
// a 64x32 sub-tile = 8x4 DMMA fragments (128 accumulator registers per thread).
//
// Shared-memory layout ("fragment order"): a 128 x 16 operand k-tile is stored as
// [k-group g = 0..3][row 0..127][4 doubles]; the A/B fragment of rows 8f..8f+7, k-group g is then the
// 256 contiguous bytes at g*4096 + f*256, read by one conflict-free LDS.64 per lane.
// Two ways of filling it: (1) Ring / mainloop_bulk: the operands already are in that order in HBM
// ("tiled layout") and one cp.async.bulk (TMA) per operand k-step lands them, mbarrier-tracked, no
// CTA-wide barrier -- what every kernel of the path uses; (2) load_operand / mainloop: per-thread
// cp.async of 16-byte pieces from a row-major matrix, one __syncthreads per k-tile -- the first
// version, kept for tools/dmma_loop.cu, which shows why (1) exists (22 % of the DMMA issue slots
// are lost to LSU contention with (2)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b7g {

#ifndef B7_BK
#define B7_BK 16
#endif
#ifndef B7_STAGES
#define B7_STAGES 4
#endif
constexpr int BM = 128, BN = 128, BK = B7_BK, STAGES = B7_STAGES, THREADS = 256;
constexpr int KG = BK / 4;                            // k-groups of 4 per staged k-tile
constexpr int OPERAND_DOUBLES = BM * BK;              // 2048 doubles = 16 KB
constexpr int STAGE_DOUBLES = 2 * OPERAND_DOUBLES;    // A then B
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;  // 128 KB

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// One 128 x 16 k-tile of a k-contiguous operand: g points at (row 0, k0), ld in doubles.
__device__ __forceinline__ void load_operand(double* s, const double* __restrict__ g, long long ld, int tid) {
  const int row = tid >> 1, h = tid & 1;
  const double* src = g + (long long)row * ld + h * 2;
  double* dst = s + row * 4 + h * 2;
#pragma unroll
  for (int g4 = 0; g4 < KG; ++g4) cp_async16(dst + g4 * (BM * 4), src + g4 * 4);
}

struct Acc {
  double c[8][4][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  }
};

// multiply-accumulate one staged k-tile: warp (wm, wn) owns rows 64*wm.., cols 32*wn..
__device__ __forceinline__ void compute_stage(const double* __restrict__ sA, const double* __restrict__ sB, int wm,
                                              int wn, int lane, Acc& acc) {
#pragma unroll
  for (int g4 = 0; g4 < KG; ++g4) {
    double a[8], b[4];
    const double* pa = sA + g4 * (BM * 4) + (64 * wm) * 4 + lane;
    const double* pb = sB + g4 * (BN * 4) + (32 * wn) * 4 + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pa[i * 32];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = pb[j * 32];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc.c[i][j][0], acc.c[i][j][1], a[i], b[j]);
  }
}

// Same for a k-tile that lies inside the diagonal 128x128 block of a lower-triangular A: rows 8f..8f+7 of the
// tile are zero right of column 8f+7, so the fragment products with k0 > row_max are skipped (the predicates
// depend on warp, fragment and k only: no divergence).  k_rel = column offset of the tile inside the block.
__device__ __forceinline__ void compute_stage_tri(const double* __restrict__ sA, const double* __restrict__ sB, int wm,
                                                  int wn, int lane, Acc& acc, int k_rel) {
#pragma unroll
  for (int g4 = 0; g4 < KG; ++g4) {
    const int k0 = k_rel + 4 * g4;
    if (k0 > 64 * wm + 63) continue;          // whole warp tile is zero from here on
    double b[4];
    const double* pa = sA + g4 * (BM * 4) + (64 * wm) * 4 + lane;
    const double* pb = sB + g4 * (BN * 4) + (32 * wn) * 4 + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = pb[j * 32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (k0 <= 64 * wm + 8 * i + 7) {
        const double a = pa[i * 32];
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc.c[i][j][0], acc.c[i][j][1], a, b[j]);
      }
    }
  }
}

// acc += A[0:128, 0:16*KT] * B[0:128, 0:16*KT]^T  (gA, gB point at the first k column)
__device__ __forceinline__ void mainloop(const double* __restrict__ gA, long long lda, const double* __restrict__ gB,
                                         long long ldb, int KT, double* smem, Acc& acc) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) {
      load_operand(smem + s * STAGE_DOUBLES, gA + s * BK, lda, tid);
      load_operand(smem + s * STAGE_DOUBLES + OPERAND_DOUBLES, gB + s * BK, ldb, tid);
    }
    cp_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    if (nxt < KT) {
      double* st = smem + (nxt % STAGES) * STAGE_DOUBLES;
      load_operand(st, gA + (long long)nxt * BK, lda, tid);
      load_operand(st + OPERAND_DOUBLES, gB + (long long)nxt * BK, ldb, tid);
    }
    cp_commit();
    const double* st = smem + (kt % STAGES) * STAGE_DOUBLES;
    compute_stage(st, st + OPERAND_DOUBLES, wm, wn, lane, acc);
  }
  cp_wait<0>();
  __syncthreads();   // smem may be reused by the caller's epilogue
}

// ---- TMA bulk copies + mbarriers (used by the posterior pass) -------------------------------------
// Operands can also live in HBM already in fragment order ("tiled layout"): tile (rb, kt) of a matrix
// with Np columns is the 2048 contiguous doubles at ((rb * (Np/16) + kt) * 2048), ordered
// [k-group][row][4].  One cp.async.bulk (SASS UBLKCP, executed by the TMA unit, no LSU issue slots)
// then moves a whole 16 KB operand k-tile, completion is signalled on an mbarrier.
constexpr int TILE_K = 16;
constexpr int TILE_DOUBLES = BM * TILE_K;   // 2048

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;\n" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "B7_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra B7_DONE;\n\t"
      "bra B7_WAIT;\n\t"
      "B7_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;\n" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// coordinates of accumulator fragment (i, j) of this lane inside the 128x128 tile
__device__ __forceinline__ int frag_row(int wm, int i, int lane) { return 64 * wm + 8 * i + (lane >> 2); }
__device__ __forceinline__ int frag_col(int wn, int j, int lane) { return 32 * wn + 8 * j + 2 * (lane & 3); }

// ---- tiled-layout addressing ---------------------------------------------------------------------
// offset of tile (rb, kt) in a matrix with kt_all = Np/16 k-tiles per row block
__host__ __device__ __forceinline__ long long tile_off(int kt_all, int rb, int kt) {
  return ((long long)rb * kt_all + kt) * TILE_DOUBLES;
}
// offset of element (row, k) inside a run of consecutive k-tiles that starts at k = 0 (row < 128)
__host__ __device__ __forceinline__ int elem_off(int row, int k) {
  return (k >> 4) * TILE_DOUBLES + ((k >> 2) & 3) * (BM * 4) + row * 4 + (k & 3);
}

// ---- bulk-copy fed main loop -----------------------------------------------------------------------
// acc += A * B^T where gA / gB point at runs of KT consecutive 16 KB tiles (tiled layout).  Thread 0 feeds
// a ring of R_STAGES stages with two cp.async.bulk per stage, R_AHEAD stages ahead of its own consumption;
// the 8 warps wait on the per-stage "full" mbarrier and release the stage through the "empty" one.
constexpr int R_STAGES = 6, R_AHEAD = 4;
constexpr int RING_BYTES = R_STAGES * 2 * TILE_DOUBLES * 8;   // 192 KB
constexpr int RING_SMEM = RING_BYTES + 1024;                  // + 2 * R_STAGES barriers (and a little slack)

struct Ring {
  double* smem;
  uint64_t *full, *empty;
  long long issued, consumed;   // stage counters since init (identical in every thread; `issued` used by thread 0)
  __device__ __forceinline__ void init(double* base) {
    smem = base;
    full = reinterpret_cast<uint64_t*>(base + R_STAGES * 2 * TILE_DOUBLES);
    empty = full + R_STAGES;
    issued = consumed = 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < R_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, THREADS / 32); }
      mbar_fence_init();
    }
    __syncthreads();
  }
  // thread 0 only
  __device__ __forceinline__ void produce(const double* a_tile, const double* b_tile) {
    const int slot = (int)(issued % R_STAGES);
    if (issued >= R_STAGES) mbar_wait(empty + slot, (unsigned)(((issued / R_STAGES) - 1) & 1));
    double* st = smem + slot * 2 * TILE_DOUBLES;
    mbar_arrive_expect_tx(full + slot, 2 * TILE_DOUBLES * 8);
    bulk_g2s(st, a_tile, TILE_DOUBLES * 8, full + slot);
    bulk_g2s(st + TILE_DOUBLES, b_tile, TILE_DOUBLES * 8, full + slot);
    ++issued;
  }
  __device__ __forceinline__ const double* wait_stage() {
    const int slot = (int)(consumed % R_STAGES);
    mbar_wait(full + slot, (unsigned)((consumed / R_STAGES) & 1));
    return smem + slot * 2 * TILE_DOUBLES;
  }
  __device__ __forceinline__ void release_stage(int lane) {
    const int slot = (int)(consumed % R_STAGES);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + slot);
    ++consumed;
  }
};

// `already` = stages the caller has produced itself before calling (to overlap its own prologue)
__device__ __forceinline__ void mainloop_bulk(Ring& ring, const double* __restrict__ gA, const double* __restrict__ gB, int KT,
                                              Acc& acc, int already = 0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 2, wn = warp & 3;
  int p = already < KT ? already : KT;
  if (tid == 0)
    for (; p < R_AHEAD && p < KT; ++p) ring.produce(gA + (long long)p * TILE_DOUBLES, gB + (long long)p * TILE_DOUBLES);
  for (int kt = 0; kt < KT; ++kt) {
    if (tid == 0 && p < KT) { ring.produce(gA + (long long)p * TILE_DOUBLES, gB + (long long)p * TILE_DOUBLES); ++p; }
    const double* st = ring.wait_stage();
    compute_stage(st, st + TILE_DOUBLES, wm, wn, lane, acc);
    ring.release_stage(lane);
  }
  __syncthreads();   // every stage consumed: the ring memory may be reused by the caller's epilogue
}

}  // namespace b7g
