// Phase-by-phase cycle stamps of the diagonal-block kernel of the Cholesky (potrf.cu: diag_kernel), the unit that sets
// the single-factor fit latency: 128 x 128 block factored, inverted and written back by one CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB7_DIAG_STAMPS -Ibot7_b200/csrc -Iinclude \
//        tools/diag_probe.cu -o tools/diag_probe -Lbot7_b200 -lbot7_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../bot7_b200'
// Prints the stamps as deltas (cycles) and the kernel time from CUDA events; checks L L^T = A and L^-1 L = I on the host.
#include <math.h>
#include <stdio.h>
#include <vector>

#include "../bot7_b200/csrc/potrf.cu"

// accuracy of the reciprocal square root used for the pivots (seed + one cubic step) against the host's 1 / sqrt
__global__ void rsq_check(const double* x, double* seed, double* full, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double piv = x[i], y = rsq_seed(piv), t = y * y, e = fma(-piv, t, 1.0), u = fma(e, 0.375, 0.5);
  seed[i] = y;
  full[i] = fma(u, y * e, y);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const int n = NBK, reps = 20;
  std::vector<double> A(n * n), At(n * n), L(n * n), Li(n * n), LiT(n * n), rhs(n);
  unsigned long long st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0 - 0.5; };
  std::vector<double> B(n * n);
  for (auto& v : B) v = rnd();
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) {
      double s = 0;
      for (int q = 0; q < n; ++q) s += B[i * n + q] * B[k * n + q];
      A[i * n + k] = s + (i == k ? 4.0 : 0.0);
    }
  for (int i = 0; i < n; ++i) rhs[i] = rnd();
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) At[elem_off(i, k)] = A[i * n + k];
  double *dA, *dA0, *dinv, *dinvT, *beta, *beta0, *logdet;
  int* info;
  CK(cudaMalloc(&dA, n * n * 8)); CK(cudaMalloc(&dA0, n * n * 8)); CK(cudaMalloc(&dinv, n * n * 8)); CK(cudaMalloc(&dinvT, n * n * 8));
  CK(cudaMalloc(&beta, n * 8)); CK(cudaMalloc(&beta0, n * 8)); CK(cudaMalloc(&logdet, 8)); CK(cudaMalloc(&info, 4));
  CK(cudaMemset(dinv, 0, n * n * 8));
  CK(cudaMemcpy(dA0, At.data(), n * n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(beta0, rhs.data(), n * 8, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f, sum = 0.f;
  for (int it = 0; it < reps; ++it) {
    CK(cudaMemcpyAsync(dA, dA0, n * n * 8, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpyAsync(beta, beta0, n * 8, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    diag_kernel<<<1, DIAG_THREADS, DIAG_SMEM>>>(dA, (long long)n * n, n, 0, dinv, (long long)n * n, beta, logdet, info, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it >= 3) { sum += ms; if (ms < best) best = ms; }
  }
  dinv_transpose_kernel<<<dim3(1, 1), 512>>>(dinv, dinvT, (long long)n * n, 0);
  CK(cudaDeviceSynchronize());
  printf("diag_kernel: %.2f us mean, %.2f us best (events, %d launches)\n", sum / (reps - 3) * 1e3, best * 1e3, reps - 3);
  long long stamps[128];
  CK(cudaMemcpyFromSymbol(stamps, b7_diag_stamps, sizeof(stamps)));
  printf("total cycles %lld; load %lld\n", stamps[35] - stamps[0], stamps[1] - stamps[0]);
  for (int p = 0; p < 8; ++p) {
    const long long start = p == 0 ? stamps[1] : stamps[4 + 3 * (p - 1)];
    if (p < 7)
      printf("  p=%d chain: factor16+inverse %5lld  wait for the update warps %5lld  look-ahead %5lld\n", p, stamps[2 + 3 * p] - start,
             stamps[3 + 3 * p] - stamps[2 + 3 * p], stamps[4 + 3 * p] - stamps[3 + 3 * p]);
    else
      printf("  p=%d chain: factor16+inverse %5lld\n", p, stamps[2 + 3 * p] - start);
  }
  printf("loop %lld; last tiles + log + assemble %lld; doubling h16 %lld %lld h32 %lld %lld h64 %lld %lld; x_j %lld; outputs %lld\n",
         stamps[26] - stamps[1], stamps[27] - stamps[26], stamps[28] - stamps[27], stamps[29] - stamps[28], stamps[30] - stamps[29],
         stamps[31] - stamps[30], stamps[32] - stamps[31], stamps[33] - stamps[32], stamps[34] - stamps[33], stamps[35] - stamps[34]);
  {
    const int m = 1 << 20;
    std::vector<double> hx(m), hs(m), hf(m);
    for (int i = 0; i < m; ++i) hx[i] = ldexp(1.0 + (rnd() + 0.5), (i % 2001) - 1000 + (i % 2));
    double *dx, *dsd, *dfl;
    CK(cudaMalloc(&dx, m * 8)); CK(cudaMalloc(&dsd, m * 8)); CK(cudaMalloc(&dfl, m * 8));
    CK(cudaMemcpy(dx, hx.data(), m * 8, cudaMemcpyHostToDevice));
    rsq_check<<<m / 256, 256>>>(dx, dsd, dfl, m);
    CK(cudaMemcpy(hs.data(), dsd, m * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hf.data(), dfl, m * 8, cudaMemcpyDeviceToHost));
    double es = 0, ef = 0;
    for (int i = 0; i < m; ++i) {
      const long double ref = 1.0L / sqrtl((long double)hx[i]);
      es = fmax(es, (double)fabsl(((long double)hs[i] - ref) / ref));
      ef = fmax(ef, (double)fabsl(((long double)hf[i] - ref) / ref));
    }
    printf("rsqrt: seed max relative error 2^%.2f, seed + cubic step %.3e (%.2f ulp) over %d values in 2^[-1000, 1001]\n", log2(es), ef,
           ef / 1.1102230246251565e-16, m);
  }
  // host check
  CK(cudaMemcpy(At.data(), dA, n * n * 8, cudaMemcpyDeviceToHost));
  std::vector<double> Di(n * n), DiT(n * n), x(n);
  CK(cudaMemcpy(Di.data(), dinv, n * n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(DiT.data(), dinvT, n * n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(x.data(), beta, n * 8, cudaMemcpyDeviceToHost));
  double ld; int inf;
  CK(cudaMemcpy(&ld, logdet, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&inf, info, 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) { L[i * n + k] = At[elem_off(i, k)]; Li[i * n + k] = Di[elem_off(i, k)]; LiT[i * n + k] = DiT[elem_off(i, k)]; }
  double e_llt = 0, e_inv = 0, e_t = 0, e_x = 0, ld_ref = 0;
  for (int i = 0; i < n; ++i) {
    ld_ref += log(L[i * n + i]);
    double xi = 0;
    for (int k = 0; k < n; ++k) {
      double s = 0, t = 0;
      for (int q = 0; q < n; ++q) { s += L[i * n + q] * L[k * n + q]; t += Li[i * n + q] * L[q * n + k]; }
      e_llt = fmax(e_llt, fabs(s - A[i * n + k]));
      e_inv = fmax(e_inv, fabs(t - (i == k ? 1.0 : 0.0)));
      e_t = fmax(e_t, fabs(LiT[i * n + k] - Li[k * n + i]));
      xi += Li[i * n + k] * rhs[k];
    }
    e_x = fmax(e_x, fabs(xi - x[i]));
  }
  printf("check: |L L^T - A| %.2e  |L^-1 L - I| %.2e  |dinvT - dinv^T| %.2e  |x - L^-1 r| %.2e  logdet %.12f (ref %.12f) info %d\n", e_llt, e_inv,
         e_t, e_x, ld, ld_ref, inf);
  return 0;
}
