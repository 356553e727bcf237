// PREPARED FOR ROUND 2 - compiled (nvcc / ptxas accept it, SASS shows UTCIMMA.2CTA) but NOT YET RUN: the round-1 GPU
// budget was spent when it was written.  Question it answers: does a CTA pair (tcgen05.mma.cta_group::2, M = 256 =
// 128 rows per CTA, N = 64, K = 32, each CTA holding its own A tile and HALF of the B tile) give the posterior
// kernel's 28-product issue pattern the same per-SM rate as cta_group::1 while each SM stages only 32 instead of
// 64 rows of the column operand (-17 % bytes per MAC, half the K* traffic)?  Part 1 validates one pair product
// bit-exactly against the CPU; part 2 times the pattern on resident operands.
//
// Layout per CTA (canonical K-major, no swizzle): A 128 rows x 32 B, element (r, k) at (k/16)*2048 + (r/8)*128 +
// (r%8)*16 + k%16; B half 32 rows x 32 B, element (n, k) at (k/16)*512 + (n/8)*128 + (n%8)*16 + k%16.
// D: CTA c holds rows 128c .. 128c+127 (TMEM lanes) x 64 columns.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(const void* smem, int lbo_bytes, int sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((s32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
template <int HINT>
__device__ __forceinline__ void mma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
#define M_(Q) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8" Q " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory")
  if (HINT == 1) M_(".collector::a::fill"); else if (HINT == 2) M_(".collector::a::use"); else if (HINT == 3) M_(".collector::a::lastuse"); else M_("");
#undef M_
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
constexpr int NS = 7, TM = 128, TNH = 32, TN = 64;   // TNH: rows of the column operand held by one CTA

template <bool PERF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe(const int8_t* __restrict__ A /* [2][NS][128 x 32 canonical] */, const int8_t* __restrict__ B /* [2][NS][32 x 32 canonical] */,
           int* __restrict__ D /* [256][64] */, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                        // NS slices x 4096 B
  uint8_t* sB = smem + NS * TM * 32;         // NS slices x 1024 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int8_t* gA = A + (size_t)rank * NS * TM * 32;
  const int8_t* gB = B + (size_t)rank * NS * TNH * 32;
  for (int e = tid; e < NS * TM * 32; e += 128) sA[e] = (uint8_t)gA[e];
  for (int e = tid; e < NS * TNH * 32; e += 128) sB[e] = (uint8_t)gB[e];
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;\n" ::"r"(1), "r"(s32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cluster_sync();                            // both CTAs have their operands, barrier and TMEM ready
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base;
  // D = S32, A = B = signed 8 bit, K-major, N = 64, M = 256 (the pair)
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  if (rank == 0 && warp == 0 && elect_one()) {
    const uint64_t da0 = make_desc(sA, TM * 16, 128), db0 = make_desc(sB, TNH * 16, 128);
    for (int it = 0; it < iters; ++it) {
      if (!PERF) {
        mma2<0>(tmem, da0, db0, idesc, 0u);  // slice 1 x slice 1 only
      } else {
#pragma unroll
        for (int p = 1; p <= NS; ++p) {
          const uint64_t da = da0 + (uint64_t)(((p - 1) * TM * 32) >> 4);
          const int nq = NS + 1 - p;
#pragma unroll
          for (int q = 1; q <= nq; ++q) {
            const uint64_t db = db0 + (uint64_t)(((q - 1) * TNH * 32) >> 4);
            const uint32_t dcol = tmem + (p + q - 2) * TN;
            if (nq == 1) mma2<0>(dcol, da, db, idesc, 1u);
            else if (q == 1) mma2<1>(dcol, da, db, idesc, 1u);
            else if (q == nq) mma2<3>(dcol, da, db, idesc, 1u);
            else mma2<2>(dcol, da, db, idesc, 1u);
          }
        }
      }
    }
    // completion lands on the barrier at the same offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(s32(&bar)), "h"((uint16_t)3) : "memory");
  }
  __syncwarp();
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c0 = 0; c0 < TN; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    if (!PERF || blockIdx.x < 2)
      for (int j = 0; j < 8; ++j) D[(size_t)(rank * TM + tid) * TN + c0 + j] = (int)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

static int canon(int r, int k, int rows) { return (k / 16) * (rows * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16); }

int main() {
  std::vector<int8_t> hA(2 * NS * TM * 32), hB(2 * NS * TNH * 32), rA(256 * 32), rB(64 * 32);
  for (int r = 0; r < 256; ++r) for (int k = 0; k < 32; ++k) rA[r * 32 + k] = (int8_t)(((r * 31 + k * 17) % 255) - 127);
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 32; ++k) rB[n * 32 + k] = (int8_t)(((n * 13 + k * 29 + 5) % 251) - 125);
  for (int c = 0; c < 2; ++c)
    for (int p = 0; p < NS; ++p) {
      for (int r = 0; r < TM; ++r) for (int k = 0; k < 32; ++k) hA[(size_t)(c * NS + p) * TM * 32 + canon(r, k, TM)] = rA[(c * TM + r) * 32 + k];
      for (int n = 0; n < TNH; ++n) for (int k = 0; k < 32; ++k) hB[(size_t)(c * NS + p) * TNH * 32 + canon(n, k, TNH)] = rB[(c * TNH + n) * 32 + k];
    }
  int8_t *dA, *dB; int* dD;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, sizeof(int) * 256 * 64));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  const int smem = NS * (TM + TNH) * 32 + 1024;
  CK(cudaFuncSetAttribute(pair_probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(pair_probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  pair_probe<false><<<2, 128, smem>>>(dA, dB, dD, 1);
  CK(cudaDeviceSynchronize());
  std::vector<int> hD(256 * 64);
  CK(cudaMemcpy(hD.data(), dD, sizeof(int) * 256 * 64, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < 256; ++i) for (int j = 0; j < 64; ++j) {
    int ref = 0;
    for (int k = 0; k < 32; ++k) ref += (int)rA[i * 32 + k] * (int)rB[j * 32 + k];
    if (ref != hD[i * 64 + j]) { if (bad < 5) printf("mismatch (%d,%d): got %d want %d\n", i, j, hD[i * 64 + j], ref); ++bad; }
  }
  printf("{\"pair_tile_mismatches\": %ld", bad);
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount / 2 * 2, iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  pair_probe<true><<<sms, 128, smem>>>(dA, dB, dD, 100); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); pair_probe<true><<<sms, 128, smem>>>(dA, dB, dD, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // one pair instruction = 256 x 64 x 32 MACs on two SMs
  printf(", \"pair_pattern_tops\": %.1f, \"ms\": %.3f, \"ctas\": %d}\n", 2.0 * 256 * 64 * 32 * 28.0 * iters * (sms / 2) / ms * 1e-9, ms, sms);
  return 0;
}
