// Dependent-chain latencies (cycles per operation, one warp) of the instructions on the critical path of the
// diagonal-block Cholesky (potrf.cu): FP64 arithmetic, conversions, the reciprocal-square-root seeds, shuffles, DMMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/lat_probe.cu -o tools/lat_probe
#include <stdio.h>
#include <cuda_runtime.h>

constexpr int N = 512;

__device__ __forceinline__ double rsq64h(double x) {
  double y;
  asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp64h(double x) {
  double y;
  asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

template <int OP>
__global__ void chain(double* out, long long* cyc, double seed, double two) {
  double x = seed + threadIdx.x * 1e-9, acc0 = 0.0, acc1 = 0.0;
  float xf = (float)seed;
  const int lane = threadIdx.x;
  __shared__ double sh[64];
  sh[lane] = (double)((lane + 1) & 31);
  sh[lane + 32] = seed;
  __syncthreads();
  int idx = lane;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, two, seed);                       // DFMA
    if (OP == 1) x = x * two;                                 // DMUL
    if (OP == 2) x = x + seed;                                // DADD
    if (OP == 3) xf = fmaf(xf, 1.0001f, 0.5f);                // FFMA
    if (OP == 4) x = (double)(float)x + seed;                 // F2F down + F2F up + DADD
    if (OP == 5) xf = rsqrtf(xf) + 1.0f;                      // MUFU.RSQ + FADD
    if (OP == 6) x = rsqrt(x) + seed;                         // library rsqrt + DADD
    if (OP == 7) x = rsq64h(x) + seed;                        // MUFU.RSQ64H + DADD
    if (OP == 8) x = rcp64h(x) + seed;                        // MUFU.RCP64H + DADD
    if (OP == 9) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);                // 64-bit shuffle (2 SHFL)
    if (OP == 10) xf = __shfl_sync(0xffffffffu, xf, (lane + 1) & 31);             // 32-bit shuffle
    if (OP == 11) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc0), "+d"(acc1) : "d"(seed), "d"(two));
    if (OP == 12) { idx = (int)sh[idx & 31]; }                // LDS.64 + F2I
    if (OP == 13) x = sqrt(x) + seed;                         // library sqrt + DADD
    if (OP == 14) x = 1.0 / x + seed;                         // library division + DADD
    if (OP == 15) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc0), "+d"(acc1) : "d"(acc0), "d"(two));   // DMMA with A depending on the result
    if (OP == 16) { x = fma(x, two, seed); acc0 = fma(acc0, two, seed); acc1 = fma(acc1, two, seed); }   // 3 independent DFMA chains
  }
  const long long t1 = clock64();
  if (lane == 0) cyc[OP] = t1 - t0;
  out[OP * 32 + lane] = x + xf + acc0 + acc1 + idx;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 32 * 8); cudaMalloc(&cyc, 32 * 8);
  const char* names[] = {"DFMA", "DMUL", "DADD", "FFMA", "F2F.f32.f64 + F2F.f64.f32 + DADD", "MUFU.RSQ + FADD", "rsqrt() + DADD",
                         "rsqrt.approx.ftz.f64 + DADD", "rcp.approx.ftz.f64 + DADD", "SHFL 64-bit", "SHFL 32-bit", "DMMA (acc chain)",
                         "LDS.64 + F2I chain", "sqrt() + DADD", "1.0 / x + DADD", "DMMA (A operand chain)", "3 independent DFMA"};
#define RUN(OP) chain<OP><<<1, 32>>>(out, cyc, 1.25, 0.75); chain<OP><<<1, 32>>>(out, cyc, 1.25, 0.75);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16)
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[32];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  for (int i = 0; i < 17; ++i) printf("%-40s %7.1f cycles per iteration\n", names[i], (double)h[i] / N);
  return 0;
}
