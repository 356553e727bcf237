// Exploration for a later round (not part of the product path): can the INT8 tcgen05 pipe carry an
// error-free-sliced (Ozaki) version of the posterior GEMM?  This probe (1) validates one
// tcgen05.mma.kind::i8 tile product (M = 128, N = 128, K = 64) against the CPU with operands written to
// shared memory in the canonical K-major no-swizzle layout (8-row x 16-byte core matrices), and
// (2) measures the issue-bound INT8 throughput of one CTA per SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: element (r, k) of an operand tile lives at (k/16)*LBO + (r/8)*SBO + (r%8)*16 + k%16
__host__ __device__ inline int canon_off(int r, int k, int rows) { return (k / 16) * (rows * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16); }

__device__ __forceinline__ uint64_t make_desc(const void* smem, int lbo_bytes, int sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((s32(smem) & 0x3FFFF) >> 4);          // start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading (K) byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // stride (M/N) byte offset
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  return d;                                             // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

__host__ __device__ inline uint32_t make_idesc_i8(int M, int N) {
  uint32_t d = 0;
  d |= 2u << 4;              // c_format = S32
  d |= 1u << 7;              // a_format = signed 8 bit
  d |= 1u << 10;             // b_format = signed 8 bit
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;                  // K-major A and B, dense, no saturate
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same MMA with an A-collector hint: 1 = fill (keep A for the next instruction), 2 = use (A comes from the collector, keep it),
// 3 = lastuse (A comes from the collector, then released).  SASS: UTCIMMA gdesc[..].A_KEEP / .A_REUSE.A_KEEP / .A_REUSE
template <int HINT>
__device__ __forceinline__ void mma_i8_coll(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (HINT == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
  else if (HINT == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
  else if (HINT == 3)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
  else mma_i8(tmem_d, da, db, idesc, accumulate);
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;\n" ::"r"(n), "r"(s32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(bar)), "r"(parity) : "memory");
}

#ifndef PN
#define PN 128
#endif
constexpr int M = 128, N = PN, K = 64;   // K in int8 elements (2 instructions of K = 32)

// one lane of a converged warp; unlike `if (lane == 0)` this lets the compiler issue UTCIMMA directly instead of wrapping
// every instruction in an ELECT / BRA.U.ANY serialisation loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred != 0;
}
#ifndef PG
#define PG 4          // MMAs per A tile in the collector experiment (the posterior kernel averages 4.5)
#endif
template <bool PERF, bool COLL>
__global__ void __launch_bounds__(128, 1) probe(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int* __restrict__ D, int iters) {
  __shared__ __align__(1024) int8_t sA[M * K];
  __shared__ __align__(1024) int8_t sB[N * K];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < M * K; e += 128) sA[e] = A[e];      // already in canonical order
  for (int e = tid; e < N * K; e += 128) sB[e] = B[e];
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = make_idesc_i8(M, N);
#ifdef PROBE_LANE0
  if (tid == 0) {
#else
  if (warp == 0 && elect_one()) {
#endif
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < K / 32; ++kk) {
        const uint64_t da = make_desc(sA + kk * 2 * (M * 16), M * 16, 128);
        const uint64_t db = make_desc(sB + kk * 2 * (N * 16), N * 16, 128);
        if (!COLL) { mma_i8(tmem, da, db, idesc, (PERF ? 1u : (uint32_t)(it > 0 || kk > 0))); continue; }
        // collector experiment: A[kk] is multiplied with PG B tiles (alternating k halves) while it stays in the collector
        const uint64_t dbx = make_desc(sB + (kk ^ 1) * 2 * (N * 16), N * 16, 128);
#pragma unroll
        for (int g = 0; g < PG; ++g) {
          const uint32_t acc = PERF ? 1u : (uint32_t)(it > 0 || kk > 0 || g > 0);
          const uint64_t dbg = (g & 1) ? dbx : db;
          if (g == 0) mma_i8_coll<1>(tmem, da, dbg, idesc, acc);
          else if (g == PG - 1) mma_i8_coll<3>(tmem, da, dbg, idesc, acc);
          else mma_i8_coll<2>(tmem, da, dbg, idesc, acc);
        }
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows), 8 columns at a time
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    if (!PERF || blockIdx.x == 0)
      for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = (int)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(N) : "memory");
}

int main() {
  std::vector<int8_t> hA(M * K), hB(N * K), rA(M * K), rB(N * K);
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) { int8_t v = (int8_t)(((r * 31 + k * 17) % 129) - 64); rA[r * K + k] = v; hA[canon_off(r, k, M)] = v; }
  for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) { int8_t v = (int8_t)(((r * 13 + k * 29 + 5) % 127) - 63); rB[r * K + k] = v; hB[canon_off(r, k, N)] = v; }
  int8_t *dA, *dB; int* dD;
  CK(cudaMalloc(&dA, M * K)); CK(cudaMalloc(&dB, N * K)); CK(cudaMalloc(&dD, sizeof(int) * M * N));
  CK(cudaMemcpy(dA, hA.data(), M * K, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), N * K, cudaMemcpyHostToDevice));
  probe<false, false><<<1, 128>>>(dA, dB, dD, 1);
  CK(cudaDeviceSynchronize());
  std::vector<int> hD(M * N);
  CK(cudaMemcpy(hD.data(), dD, sizeof(int) * M * N, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
    int ref = 0;
    for (int k = 0; k < K; ++k) ref += (int)rA[i * K + k] * (int)rB[j * K + k];
    if (ref != hD[i * N + j]) { if (bad < 5) printf("mismatch (%d,%d): got %d want %d\n", i, j, hD[i * N + j], ref); ++bad; }
  }
  printf("{\"i8_mma_tile_mismatches\": %ld", bad);
  // collector run: sum over kk, g of A[:, kk half] * B[:, (kk ^ (g & 1)) half]^T
  probe<false, true><<<1, 128>>>(dA, dB, dD, 1);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hD.data(), dD, sizeof(int) * M * N, cudaMemcpyDeviceToHost));
  long badc = 0;
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
    int ref = 0;
    for (int kk = 0; kk < 2; ++kk) for (int g = 0; g < PG; ++g) {
      const int kb = (kk ^ (g & 1)) * 32;
      for (int k = 0; k < 32; ++k) ref += (int)rA[i * K + kk * 32 + k] * (int)rB[j * K + kb + k];
    }
    if (ref != hD[i * N + j]) { if (badc < 5) printf("collector mismatch (%d,%d): got %d want %d\n", i, j, hD[i * N + j], ref); ++badc; }
  }
  printf(", \"collector_mismatches\": %ld", badc);
  // throughput: one CTA per SM, back-to-back accumulating MMAs on resident operands
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<true, false><<<p.multiProcessorCount, 128>>>(dA, dB, dD, 100); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); probe<true, false><<<p.multiProcessorCount, 128>>>(dA, dB, dD, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 2.0 * M * N * K * (double)iters * p.multiProcessorCount;
  printf(", \"n\": %d, \"i8_mma_tops\": %.1f, \"ms\": %.3f", N, ops / ms * 1e-9, ms);
  iters /= PG;
  probe<true, true><<<p.multiProcessorCount, 128>>>(dA, dB, dD, 100); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); probe<true, true><<<p.multiProcessorCount, 128>>>(dA, dB, dD, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms, e0, e1);
  ops = 2.0 * M * N * K * (double)iters * PG * p.multiProcessorCount;
  printf(", \"collector_group\": %d, \"i8_mma_collector_tops\": %.1f, \"collector_ms\": %.3f}\n", PG, ops / ms * 1e-9, ms);
  return 0;
}
