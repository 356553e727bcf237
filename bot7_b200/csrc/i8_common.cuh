// Shared pieces of the INT8 (error-free sliced) tensor paths: digit extraction and the tcgen05 wrappers used by
// posterior_i8.cu, potrf_i8.cu and trtri_i8.cu.
#pragma once
#include <stdint.h>

#include "b7_internal.h"
#include "gemm_tile.cuh"

namespace b7i8 {

using b7g::smem_u32;
constexpr int NS = B7_I8_SLICES;           // slices per operand (7)

// 7 signed digits of t in (-1, 1):  t = sum_p d_p 2^-(8p-2) + O(2^-55), d_1 in [-64, 64], the others radix 256 in
// [-128, 127].  x = rint(t 2^54) is exact in int64; adding 128 at the six low byte positions turns the balanced
// digits into the plain bytes of the sum (the carries are the 64-bit add's own), and xor 0x80 maps byte b to
// the int8 b - 128.  Result: byte k (k = 0..5) = digit 7 - k, bits 48.. = d_1 (its low byte is the int8).
__device__ __forceinline__ unsigned long long digit_bytes(double t) {
  const long long x = __double2ll_rn(t * 18014398509481984.0);   // 2^54
  return (unsigned long long)(x + 0x0000808080808080LL) ^ 0x0000808080808080ULL;
}

// same for a value that already carries the factor 2^54
__device__ __forceinline__ unsigned long long digit_bytes_scaled(double t54) {
  const long long x = __double2ll_rn(t54);
  return (unsigned long long)(x + 0x0000808080808080LL) ^ 0x0000808080808080ULL;
}

// digit bytes of 4 consecutive k -> one 32-bit word per slice (byte j = element j): two 4 x 4 byte transposes
__device__ __forceinline__ void pack4(const unsigned long long (&z)[4], uint32_t (&w)[NS]) {
  const uint32_t l0 = (uint32_t)z[0], l1 = (uint32_t)z[1], l2 = (uint32_t)z[2], l3 = (uint32_t)z[3];
  const uint32_t h0 = (uint32_t)(z[0] >> 32), h1 = (uint32_t)(z[1] >> 32), h2 = (uint32_t)(z[2] >> 32), h3 = (uint32_t)(z[3] >> 32);
  const uint32_t la = __byte_perm(l0, l1, 0x5140), lb = __byte_perm(l0, l1, 0x7362);   // [0.b0 1.b0 0.b1 1.b1], [0.b2 1.b2 0.b3 1.b3]
  const uint32_t lc = __byte_perm(l2, l3, 0x5140), ld = __byte_perm(l2, l3, 0x7362);
  const uint32_t ha = __byte_perm(h0, h1, 0x5140), hb = __byte_perm(h0, h1, 0x7362);
  const uint32_t hc = __byte_perm(h2, h3, 0x5140), hd = __byte_perm(h2, h3, 0x7362);
  w[6] = __byte_perm(la, lc, 0x5410);   // digit 7 = byte 0 of the low words
  w[5] = __byte_perm(la, lc, 0x7632);
  w[4] = __byte_perm(lb, ld, 0x5410);
  w[3] = __byte_perm(lb, ld, 0x7632);
  w[2] = __byte_perm(ha, hc, 0x5410);   // digit 3 = byte 0 of the high words
  w[1] = __byte_perm(ha, hc, 0x7632);
  w[0] = __byte_perm(hb, hd, 0x5410);   // digit 1 = byte 2 of the high words (|d_1| <= 64)
}

// ---- tcgen05 helpers -----------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t umma_desc(const void* smem, int lbo_bytes, int sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // K-direction stride between 16-byte chunks
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // M/N-direction stride between 8-row groups
  d |= (uint64_t)1 << 46;                             // sm_100 descriptor version; SWIZZLE_NONE
  return d;
}
// HINT: 0 = plain, 1 = keep A in the collector (fill), 2 = A from the collector and keep it (use),
// 3 = A from the collector, then release it (lastuse).  SASS: UTCIMMA gdesc[..].A_KEEP / .A_REUSE.A_KEEP / .A_REUSE
template <int HINT>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
#define B7_UMMA_I8(QUAL)                                                                                            \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8" QUAL " [%0], %1, %2, %3, p;\n\t}\n" \
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory")
  if (HINT == 1) B7_UMMA_I8(".collector::a::fill");
  else if (HINT == 2) B7_UMMA_I8(".collector::a::use");
  else if (HINT == 3) B7_UMMA_I8(".collector::a::lastuse");
  else B7_UMMA_I8("");
#undef B7_UMMA_I8
}
// one lane of a converged warp (all 32 lanes must call it)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ---- CTA pairs (tcgen05 cta_group::2) --------------------------------------------------------------------
// One instruction of the leader CTA drives the tensor cores of both SMs of a cluster of two: M = 256 (CTA c owns rows
// 128 c .. 128 c + 127 of A and of D, in its own shared memory / TMEM), B is split by rows of the N dimension (CTA c
// holds columns N/2 c .. of the output), and both shared memories are read at the same offsets.
template <int HINT>
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
#define B7_UMMA_I8P(QUAL)                                                                                           \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8" QUAL " [%0], %1, %2, %3, p;\n\t}\n" \
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory")
  if (HINT == 1) B7_UMMA_I8P(".collector::a::fill");
  else if (HINT == 2) B7_UMMA_I8P(".collector::a::use");
  else if (HINT == 3) B7_UMMA_I8P(".collector::a::lastuse");
  else B7_UMMA_I8P("");
#undef B7_UMMA_I8P
}
// completion of every MMA issued so far by this thread -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// one arrival on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Default semantics
// (release at CTA scope), the form CUTLASS's ClusterBarrier::arrive(cta_id) uses: `.release.cluster` makes ptxas
// emit MEMBAR.ALL.GPU + ERRBAR in front of every arrive (measured: 43 % of the relay warp's time, ncu pair_r02).
// What crosses the pair here is ordered by the barriers themselves: TMA writes complete on the peer's own barrier
// before the relay arrives, and the tensor core reads them through the async proxy after the leader's wait.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(
                   smem_u32(bar)),
               "r"(cta)
               : "memory");
}
// wait (arrivals may come from the peer CTA) with a watchdog: a protocol error turns into a trap after ~2 s
// instead of a hung GPU
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, unsigned parity) {
  uint32_t ok = 0;
  unsigned long long t0 = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (ok) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    if (t0 == 0) t0 = t;
    else if (t - t0 > 2000000000ULL) __trap();
  }
}

// L2 eviction-priority hints for the bulk copies: the slices of L^-1 are re-read by every CTA pair for the whole
// launch (evict_last), a candidate tile's K* slices are streamed (evict_first / normal)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// tcgen05.ld without the wait: several loads can be in flight before one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_async(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr));
}
// wait for every tcgen05.ld issued so far; the loaded registers are passed through empty asm statements so that the
// compiler cannot move their first use above the wait
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_pin(uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
}

// ---- slicing kernels: 512 threads = 128 rows (or columns) x 4 interleaved k quarters --------------------------
constexpr int SLICE_THREADS = 512;

// power-of-two scale of one row from the partial maxima of its four quarters (NaN when any entry was not finite:
// a failed pivot must poison what it touches, as it does on the fp64 path).  s_red: 512 doubles of shared memory.
__device__ __forceinline__ double row_scale(double mx, bool bad, double* s_red, int row, int q) {
  s_red[q * 128 + row] = bad ? __longlong_as_double(0x7ff8000000000000LL) : mx;
  __syncthreads();
  double m = 0.0;
  bool b = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double v = s_red[j * 128 + row];
    b |= (v != v);
    m = fmax(m, v);
  }
  int e = 0;
  frexp(m, &e);                                    // 2^e > m
  return b ? __longlong_as_double(0x7ff8000000000000LL) : (m > 0.0 ? ldexp(1.0, e) : 1.0);
}

// 16 consecutive entries of one row in the tiled fp64 layout (p points at elem_off(row, 16 c)): max / digits
__device__ __forceinline__ void chunk_max(const double* p, double& mx, bool& bad) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const double4 v = *reinterpret_cast<const double4*>(p + g * (128 * 4));
    bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
    mx = fmax(mx, fmax(fmax(fabs(v.x), fabs(v.y)), fmax(fabs(v.z), fabs(v.w))));
  }
}
__device__ __forceinline__ void chunk_digits(const double* p, double inv, uint32_t (&pk)[4][NS]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const double4 v = *reinterpret_cast<const double4*>(p + g * (128 * 4));
    const unsigned long long z[4] = {digit_bytes(v.x * inv), digit_bytes(v.y * inv), digit_bytes(v.z * inv), digit_bytes(v.w * inv)};
    pack4(z, pk[g]);
  }
}

// ---- the 128 x 64 x 64 stage shared by the factorisation kernels (potrf_i8.cu, trtri_i8.cu) ----------------
// A stage = 7 slices x 128 rows x 64 k-bytes of the row operand followed by 7 x 64 x 64 of the column operand, both
// in the UMMA canonical K-major no-swizzle layout [slice][16-byte k chunk][row][16].
namespace gemm {
constexpr int TM = 128, TN = 64, KB = 64, KC = KB / 16;
constexpr int A_STAGE = NS * TM * KB;      // 57344 B
constexpr int B_STAGE = NS * TN * KB;      // 28672 B
constexpr int STAGE = A_STAGE + B_STAGE;   // 86016 B
constexpr int NSTAGE = 2;
constexpr int THREADS = 192;               // warps 0-3 epilogue, 4 producer, 5 MMA issue
constexpr int SMEM = NSTAGE * STAGE + 1024;

// instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N = 64, M = 128
__device__ __forceinline__ uint32_t idesc_128x64() {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// the 56 MMAs of one stage (call from ONE elected lane): slice p of the row operand meets slices 1 .. 8 - p of the
// column operand from the collector; class p + q accumulates in TMEM columns (p + q - 2) * 64
__device__ __forceinline__ void mma_stage(const uint8_t* stage, uint32_t tmem, uint32_t idesc, bool first) {
  const uint64_t da0 = umma_desc(stage, TM * 16, 128), db0 = umma_desc(stage + A_STAGE, TN * 16, 128);
#pragma unroll
  for (int k2 = 0; k2 < KB / 32; ++k2) {
#pragma unroll
    for (int p = 1; p <= NS; ++p) {
      const uint64_t da = da0 + (uint64_t)(((p - 1) * (KC * TM * 16) + k2 * (2 * TM * 16)) >> 4);
      const uint32_t acc = (first && k2 == 0 && p == 1) ? 0u : 1u;
      const int nq = NS + 1 - p;
#pragma unroll
      for (int q = 1; q <= nq; ++q) {
        const uint64_t db = db0 + (uint64_t)(((q - 1) * (KC * TN * 16) + k2 * (2 * TN * 16)) >> 4);
        const uint32_t dcol = tmem + (uint32_t)((p + q - 2) * TN);
        if (nq == 1) umma_i8<0>(dcol, da, db, idesc, acc);
        else if (q == 1) umma_i8<1>(dcol, da, db, idesc, acc);
        else if (q == nq) umma_i8<3>(dcol, da, db, idesc, acc);
        else umma_i8<2>(dcol, da, db, idesc, acc);
      }
    }
  }
}

// epilogue warp `warp` (0-3), thread = TMEM lane: v[c] = sum_w 2^(4-8w) S_w[lane][c] for the 64 columns
__device__ __forceinline__ void drain_classes(uint32_t tmem, int warp, double (&v)[TN]) {
#pragma unroll
  for (int c = 0; c < TN; ++c) v[c] = 0.0;
#pragma unroll
  for (int wc = 2; wc <= NS + 1; ++wc) {
    const double wt = ldexp(1.0, 4 - 8 * wc);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t dv[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((wc - 2) * TN + hh * 32), dv);
#pragma unroll
      for (int c = 0; c < 32; ++c) v[hh * 32 + c] = fma((double)(int)dv[c], wt, v[hh * 32 + c]);
    }
  }
}
}  // namespace gemm

}  // namespace b7i8
