// Where does the tile engine lose DMMA issue slots?  Runs compute_stage() of gemm_tile.cuh on
// resident shared memory (no global traffic) with / without barriers and cp.async traffic.
#include <cstdio>
#include <cuda_runtime.h>
#include "../bot7_b200/csrc/gemm_tile.cuh"
using namespace b7g;

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(double* out, const double* src, int iters) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 2, wn = warp & 3;
  for (int e = tid; e < STAGES * STAGE_DOUBLES; e += THREADS) smem[e] = 1e-3 * (e % 97);
  __syncthreads();
  Acc acc; acc.zero();
  for (int it = 0; it < iters; ++it) {
    if (MODE >= 1) __syncthreads();
    if (MODE >= 2) {
      double* st = smem + ((it + STAGES - 1) % STAGES) * STAGE_DOUBLES;
      load_operand(st, src + (size_t)(blockIdx.x % 8) * 128 * 4096 + (it % 256) * BK, 4096, tid);
      load_operand(st + OPERAND_DOUBLES, src + (size_t)(8 + blockIdx.x) * 128 * 4096 + (it % 256) * BK, 4096, tid);
      cp_commit();
      cp_wait<STAGES - 2>();
    }
    const double* st = smem + (it % STAGES) * STAGE_DOUBLES;
    compute_stage(st, st + OPERAND_DOUBLES, wm, wn, lane, acc);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc.c[i][j][0] + acc.c[i][j][1];
  out[blockIdx.x * THREADS + tid] = s;
}

template <int MODE> float run(double* out, const double* src, int iters) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, THREADS, SMEM_BYTES>>>(out, src, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148, THREADS, SMEM_BYTES>>>(out, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  double *out, *src; cudaMalloc(&out, 148 * THREADS * 8); cudaMalloc(&src, (size_t)160 * 128 * 4096 * 8);
  cudaMemset(src, 0, (size_t)160 * 128 * 4096 * 8);
  int iters = 4000;
  double flops = 2.0 * 128 * 128 * BK * iters * 148;
  float m0 = run<0>(out, src, iters), m1 = run<1>(out, src, iters), m2 = run<2>(out, src, iters);
  printf("BK=%d STAGES=%d  smem-only %.2f TF | +barrier %.2f TF | +cp.async stream %.2f TF  (%s)\n", BK, STAGES,
         flops / m0 * 1e-9, flops / m1 * 1e-9, flops / m2 * 1e-9, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
