// Sustained (power-capped) INT8 MMA rate for different issue orders of the 28-product slice pattern of
// posterior_i8.cu (development aid).  Operands are resident in shared memory, so the only variables are the
// order of the MMAs, the accumulator they target and whether the A tile is held in the collector:
//   mode 0: p-major, A collector, class accumulators p + q     (the kernel's order)
//   mode 1: p-major, no collector
//   mode 2: class-major (all products of one accumulator back to back), no collector
//   mode 3: p-major, A collector, ONE accumulator (no TMEM rotation)
// Each mode runs ~1.5 s in 40 ms launches; the first and the last launches show what the power cap does.
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
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
#define M_(Q) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8" Q " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory")
  if (HINT == 1) M_(".collector::a::fill"); else if (HINT == 2) M_(".collector::a::use"); else if (HINT == 3) M_(".collector::a::lastuse"); else M_("");
#undef M_
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred != 0;
}
constexpr int NS = 7, TM = 128, TN = 64;
template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(int iters, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                      // 7 slices x 128 rows x 32 B
  uint8_t* sB = smem + NS * TM * 32;       // 7 slices x 64 rows x 32 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < NS * (TM + TN) * 32; e += 128) smem[e] = (uint8_t)((e * 37 + 11) & 0xff);
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;\n" ::"r"(1), "r"(s32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  if (warp == 0 && elect_one()) {
    const uint64_t da0 = make_desc(sA, TM * 16, 128), db0 = make_desc(sB, TN * 16, 128);
    for (int it = 0; it < iters; ++it) {
      if (MODE == 2) {
#pragma unroll
        for (int w = 2; w <= NS + 1; ++w)
#pragma unroll
          for (int p = 1; p < w; ++p) {
            const int q = w - p;
            mma<0>(tmem + (w - 2) * TN, da0 + (uint64_t)(((p - 1) * TM * 32) >> 4), db0 + (uint64_t)(((q - 1) * TN * 32) >> 4), idesc);
          }
      } else {
#pragma unroll
        for (int p = 1; p <= NS; ++p) {
          const uint64_t da = da0 + (uint64_t)(((p - 1) * TM * 32) >> 4);
          const int nq = NS + 1 - p;
#pragma unroll
          for (int q = 1; q <= nq; ++q) {
            const uint64_t db = db0 + (uint64_t)(((q - 1) * TN * 32) >> 4);
            const uint32_t dcol = tmem + (MODE == 3 ? 0 : (p + q - 2) * TN);
            if (MODE == 1 || nq == 1) mma<0>(dcol, da, db, idesc);
            else if (q == 1) mma<1>(dcol, da, db, idesc);
            else if (q == nq) mma<3>(dcol, da, db, idesc);
            else mma<2>(dcol, da, db, idesc);
          }
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(&bar)) : "memory");
  }
  __syncwarp();
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(v) : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  if (v == 0x12345678u) sink[0] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}
template <int MODE>
void run(int sms, int* sink) {
  const int smem = NS * (TM + TN) * 32 + 1024;
  CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 60000;     // 28 MMAs each: ~40 ms
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<sms, 128, smem>>>(1000, sink); CK(cudaDeviceSynchronize());
  std::vector<double> tops;
  for (int rep = 0; rep < 36; ++rep) {
    cudaEventRecord(e0); probe<MODE><<<sms, 128, smem>>>(iters, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    tops.push_back(2.0 * TM * TN * 32 * 28.0 * iters * sms / ms * 1e-9);
  }
  double first = (tops[0] + tops[1]) / 2, last = 0; for (int i = 26; i < 36; ++i) last += tops[i] / 10;
  printf("{\"mode\": %d, \"tops_first_80ms\": %.1f, \"tops_sustained\": %.1f}\n", MODE, first, last);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int* sink; CK(cudaMalloc(&sink, 4));
  run<0>(p.multiProcessorCount, sink); run<1>(p.multiProcessorCount, sink); run<2>(p.multiProcessorCount, sink); run<3>(p.multiProcessorCount, sink);
  return 0;
}
