// FP64 peak probe for B200 (sm_100a): raw DMMA.8x8x4 issue rate, raw DFMA rate, cuBLAS DGEMM.
// Output: one JSON object on stdout. Used to fix the FP64 roofline denominator.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_loop(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_loop(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d", p.name, sms, p.l2CacheSize);
  const int iters = 20000;
  // DMMA: vary warps/SM and accumulators/warp
  int wps[] = {4, 8, 16, 32};
  for (int w : wps) {
    int threads = (w >= 8 ? 256 : w * 32), blocks_per_sm = (w * 32) / threads;
    int blocks = sms * blocks_per_sm;
    float ms8 = time_ms([&] { dmma_loop<8><<<blocks, threads>>>(out, iters, 1.0, 1e-3); }, 5);
    float ms32 = time_ms([&] { dmma_loop<32><<<blocks, threads>>>(out, iters / 4, 1.0, 1e-3); }, 5);
    double fl8 = 512.0 * 8 * iters * (double)w * sms, fl32 = 512.0 * 32 * (iters / 4) * (double)w * sms;
    printf(", \"dmma_w%d_acc8_tflops\": %.3f, \"dmma_w%d_acc32_tflops\": %.3f", w, fl8 / ms8 * 1e-9, w, fl32 / ms32 * 1e-9);
  }
  for (int w : wps) {
    int threads = (w >= 8 ? 256 : w * 32), blocks_per_sm = (w * 32) / threads;
    int blocks = sms * blocks_per_sm;
    float ms = time_ms([&] { dfma_loop<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-3); }, 5);
    double fl = 2.0 * 16 * iters * 32.0 * w * sms;
    printf(", \"dfma_w%d_tflops\": %.3f", w, fl / ms * 1e-9);
  }
  CK(cudaGetLastError());
  // cuBLAS DGEMM
  cublasHandle_t h; cublasCreate(&h);
  int ns[] = {4096, 8192};
  for (int n : ns) {
    double *A, *B, *C; size_t bytes = sizeof(double) * n * n;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    std::vector<double> hA((size_t)n * n);
    for (size_t i = 0; i < hA.size(); ++i) hA[i] = (double)((i * 2654435761u) % 1000) * 1e-3 - 0.5;
    CK(cudaMemcpy(A, hA.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hA.data(), bytes, cudaMemcpyHostToDevice));
    double one = 1.0, zero = 0.0;
    float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
    printf(", \"cublas_dgemm_tn_%d_tflops\": %.3f", n, 2.0 * n * n * (double)n / ms * 1e-9);
    ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
    printf(", \"cublas_dgemm_nn_%d_tflops\": %.3f", n, 2.0 * n * n * (double)n / ms * 1e-9);
    if (n == 8192) {  // sustained: back-to-back for ~3 s
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      int reps = 60; cudaEventRecord(e0);
      for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1);
      printf(", \"cublas_dgemm_tn_8192_sustained_tflops\": %.3f", 2.0 * n * n * (double)n * reps / t * 1e-9);
    }
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf(", \"sm_clock_khz_attr\": %d}\n", clk);
  return 0;
}
