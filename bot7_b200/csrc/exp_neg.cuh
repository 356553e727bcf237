// Table-driven fp64 exp for non-positive arguments, shared by the covariance builders (cov.cu, posterior_i8.cu).
#pragma once
#include <cuda_runtime.h>

namespace {

// exp(x) for x <= 0, table driven: x = n ln2/64 + r, exp(x) = 2^(n>>6) * 2^((n&63)/64) * (1 + p(r)), |r| <= ln2/128,
// p of degree 5 (truncation 3e-17).  ~12 FP64 instructions instead of ~30 for exp(): this kernel is bound by
// the FP64 issue rate, not by the 8 bytes it writes per entry.  <= 1 ulp measured against glibc on 1e7 points;
// results below the normal range (x < -708) are flushed to 0.
__constant__ double c_exp_tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tab) {
  if (!(x > -708.0)) return (x != x) ? x : 0.0;
  const double t = fma(x, 92.33248261689366, 6755399441055744.0);   // 64/ln2, 1.5*2^52: n in the low word
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -0.010830424696249145, x);          // ln2/64 high part
  r = fma(nf, -(3.623510646634843e-19), r);                 // ln2/64 low part
  double p = fma(r, 8.33333333333333322e-03, 4.16666666666666644e-02);
  p = fma(p, r, 1.66666666666666657e-01);
  p = fma(p, r, 0.5);
  p = fma(p * r, r, r);
  const double T = tab[n & 63];
  const double v = fma(T, p, T);
  return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}

// The same function with its six non-trivial constants held in registers (made opaque once per kernel) and the range
// check as a select at the end.  As written above, the compiler re-materialises every 64-bit constant in front of its DFMA
// (two moves each) and wraps the call in BSSY / BRA / BSYNC: 16 + 5 of the 85 instructions per K* entry of
// cov_slices_kernel, which is issue-bound (ncu covslices_r02: FP64 pipe 32 % active, issue slots 71 % busy).  Same FP64
// operations in the same order: bit-identical results for every argument.
struct ExpNegK {
  double l2e, ln2h, ln2l, c5, c4, c3;
  __device__ __forceinline__ void init() {
    l2e = 92.33248261689366; ln2h = -0.010830424696249145; ln2l = -(3.623510646634843e-19);
    c5 = 8.33333333333333322e-03; c4 = 4.16666666666666644e-02; c3 = 1.66666666666666657e-01;
    asm volatile("" : "+d"(l2e), "+d"(ln2h), "+d"(ln2l), "+d"(c5), "+d"(c4), "+d"(c3));
  }
};

__device__ __forceinline__ double exp_neg_k(double x, const double* __restrict__ tab, const ExpNegK& K) {
  const double t = fma(x, K.l2e, 6755399441055744.0);
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, K.ln2h, x);
  r = fma(nf, K.ln2l, r);
  double p = fma(r, K.c5, K.c4);
  p = fma(p, r, K.c3);
  p = fma(p, r, 0.5);
  p = fma(p * r, r, r);
  const double T = tab[n & 63];
  const double v = fma(T, p, T);
  const double res = __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
  return x > -708.0 ? res : ((x != x) ? x : 0.0);
}

}  // namespace
