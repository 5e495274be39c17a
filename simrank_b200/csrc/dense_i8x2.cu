// Paired-SM tensor-core half-product for the dense SimRank chain on sm_100a.
//
//   D[j, r] = sum_k A8[j, k] * V[r, k]          j < M, r < R, k < K
//
// A8 is the 0/1 adjacency pattern (the M-side operand: 256 rows per CTA pair), V a non-negative
// matrix held as NS uint8 fixed-point planes (V[r,k] ~= q[r,k] * bound(r) / 256^NS).  The NS planes
// of one block of RT rows of V are CONCATENATED on the N side of a single
// tcgen05.mma.cta_group::2.kind::i8 (u8 x u8 -> s32, M = 256, N = NS*RT, K = 32): column
// p*RT + c of the accumulator is plane p of row r0 + c.  Products and K-long sums are exact
// integers; the epilogue recombines the planes in 64-bit integers.  See DESIGN.md "K1/K2".
//
// Why this shape.  With M = 128 per CTA the int8 pipe consumes its N-side operand at 64 B/clk and
// its M-side operand at 8192/N B/clk from shared memory while TMA refills the same bytes: the
// single-CTA kernel (dense_i8.cu) needed 187 B/clk against a 128 B/clk port and sat at 60-65 % of
// the pipe.  A CTA pair splits the N-side operand between the two SMs (each holds N/2 rows) and
// one instruction covers all planes, so the M-side tile is read once per K-step instead of NS
// times: 128 B/clk (NS = 2, N = 256) / 149 B/clk (NS = 3, N = 192).  The accumulator needs only
// N <= 256 TMEM columns, so it is double-buffered and the epilogue of tile t overlaps the
// mainloop of tile t+1.  The result is row-major in j with r along the TMEM columns, so no
// transposed store is needed between the two half-products.
//
// Roles (256 threads per CTA, cluster of 2 CTAs, persistent over pair tiles of 256 x RT):
//   warp 0    TMA producer (both CTAs): its 128 rows of A8 + its N/2 rows of the planes per K-block
//             into a kStages-deep 128B-swizzled ring; every copy signals the LEADER's full barrier
//   warp 1    MMA issuer (leader CTA only): 4 tcgen05.mma per K-block; tcgen05.commit multicasts
//             the "slot free" / "accumulator full" arrivals to both CTAs
//   warp 2    TMEM allocator (512 columns = 2 accumulator buffers)
//   warps 4-7 epilogue: tcgen05.ld -> exact recombination -> fused epilogue (MID / FINAL / COUNTS)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace srk {
namespace x2 {

constexpr int BMC = 128;        // A8 rows per CTA (= TMEM lanes)
constexpr int BK = 128;         // bytes of K per pipeline stage (= one 128B swizzle atom)
constexpr int UMMA_K = 32;      // K per tcgen05.mma for 8-bit operands
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;
constexpr int kSmemLimit = 232448;
constexpr int kAccStride = 256; // TMEM columns per accumulator buffer
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address

template <int NS>
struct Cfg {
  static constexpr int RT = NS == 1 ? 256 : NS == 2 ? 128 : 64;     // rows of V per tile
  static constexpr int N = NS * RT;                                 // UMMA N (multiple of 32)
  static constexpr int NH = N / 2;                                  // N-side rows held by each CTA
  static constexpr int BR = NS == 3 ? 32 : (NS == 4 ? 64 : 128);    // rows per TMA box (divides NH and RT)
  static constexpr int kBoxes = NH / BR;
  static constexpr int kStageBytes = BMC * BK + NH * BK;
  static constexpr int kColfacBytes = 2 * 2 * RT * 8;               // 2 buffers x 2 vectors
  static constexpr int kMirrorRow = 18;                             // doubles per staged row (16 + pad: 144 B)
  static constexpr int kMirrorBytes = 4 * 32 * kMirrorRow * 8;      // one 32 x 16 tile per epilogue warp
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - 256 - kColfacBytes - kMirrorBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kColfacBytes + 256 + kMirrorBytes;
  static_assert(N % 32 == 0 && N <= 256, "invalid UMMA N for cta_group::2 kind::i8");
  static_assert(NH % BR == 0 && RT % BR == 0, "box rows must divide the CTA half and the plane block");
  static_assert(kStages >= 3, "not enough shared memory for a pipeline");
};

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// TMA loads of a CTA pair: the data lands in THIS CTA's shared memory, the transaction bytes are
// counted on the leader CTA's barrier (rank bit cleared).
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, uint64_t pol, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, uint64_t pol, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, uint64_t pol, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
// L2 eviction policies.  The operand panels are re-read by the other CTA pairs of a wave (keep:
// evict_last); everything the epilogues touch is used exactly once (evict_first), so that 60 MB of
// result traffic per wave does not push the panels out of the 126 MB L2.
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ld_stream_f64x2(const double2* a, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* a, uint64_t pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u32x4(const uint4* a, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_stream_f64x2(double2* a, double x, double y, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(a), "d"(x), "d"(y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_f64(double* a, double x, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(x), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_u32x4(uint4* a, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs] * B[smem of both CTAs]^T, u8 x u8 -> s32
__device__ __forceinline__ void mma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives (in both CTAs of the pair) on the barrier at this offset once every tcgen05.mma issued
// so far by this thread has completed.
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void red_add_gpu(unsigned int* a) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(a) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* a) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// K-major operand tile, rows of 128 B, 128B swizzle: 8-row groups are 1024 B apart (SBO),
// LBO unused for swizzled K-major layouts (encoded 1), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::i8 instruction descriptor: D = s32, A = B = u8, both K-major, M = 256 (pair), N
template <int N>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// exact value of sum_p acc_p * 256^(NS-1-p) (< 2^(8 NS + 22) for every supported K)
template <int NS>
__device__ __forceinline__ unsigned long long combine(const uint32_t (&a)[NS][16], int x) {
  unsigned long long v = 0;
#pragma unroll
  for (int s = 0; s < NS; ++s) v = (v << 8) + (unsigned long long)a[s][x];
  return v;
}

// The planes of U = A S_off (MID output, FINAL input) are cut with POWER-OF-TWO row bounds: the
// caller's bound b is rounded up to 2^f.  That costs at most one bit of U's resolution and turns
// every per-element scaling that involves the bound into integer shifts / exponent arithmetic: in
// situ (tensor pipe busy) a DMUL/DADD/DSETP costs ~60 issue cycles per warp, so the epilogues are
// written around "as few FP64-pipe instructions per element as possible" (MID 1, FINAL 3).
constexpr int kNoBound = -(1 << 30);
template <int NS>
__device__ __forceinline__ int pow2_exponent(double b) {
  if (!(b > 0.0)) return kNoBound;                        // all-zero row
  const long long bits = __double_as_longlong(b);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  if (bits & 0xfffffffffffffll) ++e;
  const int lo = 8 * NS - 40;                             // keeps counts (< 2^22) << (8 NS - f) inside 62 bits
  return e < lo ? lo : (e > 1000 ? 1000 : e);
}
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }

// MID: q = rn(D * alpha * 2^-f), saturated to 256^NS - 1.  One FP64-pipe multiply; the power of
// two is applied to the exponent field of the converted integer.
template <int NS>
__device__ __forceinline__ uint32_t mid_quant(unsigned long long D, double alpha, int f) {
  long long bits = __double_as_longlong((double)(long long)D);
  bits -= (long long)f << 52;
  const double d = D ? __longlong_as_double(bits) : 0.0;
  uint32_t q = __double2uint_rn(d * alpha);               // saturating; NaN -> 0
  if (NS < 4) q = min(q, (uint32_t)((1ull << (8 * NS)) - 1ull));
  return q;
}

// FINAL: x = coef * g_a[j] * g_v[r] * (D * 2^(f_r - 8 NS) + count).  With s = max(8 NS - f_r, 0),
// dl = max(f_r - 8 NS, 0), T = (D << dl) + (count << s) (an exact integer < 2^62) and the two
// positive scale factors cf = coef * g_v[r] * 2^-s and gj = g_a[j] this is x = T * cf * gj.
//
// The product is formed with INTEGER instructions.  FP64-pipe instructions (DMUL, DADD, I2F.F64)
// issued while the tensor pipe is saturated take ~90 cycles each (ncu: 40 % of the epilogue's
// stall samples were math-pipe throttle with 4 of them per element) and made the FINAL epilogue
// longer than the mainloop it has to hide behind.  Every factor is held as a 64-bit mantissa with
// bit 63 set plus a binary exponent, x = m / 2^63 * 2^e; two mul.hi.u64 give the top 64 bits of
// T * cf * gj (relative truncation error < 2^-60), which are rounded to the 53-bit mantissa of the
// result: within 0.51 ulp of the exactly rounded product, where the FP64 chain rounded twice.
__device__ __forceinline__ void split_f64(double x, unsigned long long& m, int& e) {
  const long long bits = __double_as_longlong(x);
  const int ex = (int)((bits >> 52) & 0x7ff);
  const bool ok = bits > 0 && ex != 0 && ex != 0x7ff;          // positive, normal, finite; else 0
  m = ok ? (((unsigned long long)bits & 0xfffffffffffffull) | (1ull << 52)) << 11 : 0ull;
  e = ex - 1023;
}
// sh = s | dl << 8 | (e_cf + 2048) << 16
__device__ __forceinline__ int pack_shift(int s, int dl, int e_cf) { return s | (dl << 8) | ((e_cf + 2048) << 16); }
__device__ __forceinline__ double final_value(unsigned long long D, uint32_t cnt, int sh, unsigned long long m_cf,
                                              unsigned long long m_gj, int e_gj) {
  const unsigned long long T = (D << ((sh >> 8) & 0xff)) + ((unsigned long long)cnt << (sh & 0xff));
  // c = cf * gj = mc / 2^63 * 2^ec
  unsigned long long mc = __umul64hi(m_cf, m_gj);               // in [2^62, 2^64) or 0
  int ec = (sh >> 16) - 2048 + e_gj;
  if ((long long)mc < 0) ++ec; else mc <<= 1;
  const int z = __clzll((long long)T) & 63;
  unsigned long long pr = __umul64hi(T << z, mc);               // T * c = pr * 2^(1 - z + ec), pr in [2^62, 2^64)
  int E = 63 - z + ec;
  if ((long long)pr < 0) ++E; else pr <<= 1;
  const int biased = E + 1023;
  unsigned long long bits = ((unsigned long long)(unsigned)(biased - 1) << 52) + (pr >> 11) + ((pr >> 10) & 1ull);
  if (T == 0ull || mc == 0ull || biased <= 0) bits = 0ull;      // zero operands, underflow
  return __longlong_as_double((long long)bits);
}
// common-neighbour counts are held as uint16 (or uint32 when a degree can reach 65535)
__device__ __forceinline__ uint32_t load_count(const void* base, int64_t idx, int c32) {
  return c32 ? reinterpret_cast<const uint32_t*>(base)[idx] : (uint32_t)reinterpret_cast<const uint16_t*>(base)[idx];
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
// High word of a non-negative double, rounded up: a 32-bit key that is monotone in the value
// (keys order like the doubles) and whose "value" (key << 32) is an upper bound within 2^-20.
__device__ __forceinline__ uint32_t hi_key(double x) { return (uint32_t)__double2hiint(x) + 1u; }
// Column maxima of a 32 x 16 block held as 16 keys per lane: after the call lane l holds the
// maximum over all 32 lanes of column col_of_lane(l) = bits 4..1 of l (both lanes of a pair hold
// the same column).  Reduce-scatter: 16 shuffles instead of 80.
__device__ __forceinline__ uint32_t column_max16(const uint32_t (&k)[16], int lane) {
  uint32_t a[8], b[4], c[2], d;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t mine = h16 ? k[i + 8] : k[i], other = h16 ? k[i] : k[i + 8];
    a[i] = max(mine, __shfl_xor_sync(0xffffffffu, other, 16));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t mine = h8 ? a[i + 4] : a[i], other = h8 ? a[i] : a[i + 4];
    b[i] = max(mine, __shfl_xor_sync(0xffffffffu, other, 8));
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t mine = h4 ? b[i + 2] : b[i], other = h4 ? b[i] : b[i + 2];
    c[i] = max(mine, __shfl_xor_sync(0xffffffffu, other, 4));
  }
  {
    const uint32_t mine = h2 ? c[1] : c[0], other = h2 ? c[0] : c[1];
    d = max(mine, __shfl_xor_sync(0xffffffffu, other, 2));
  }
  return max(d, __shfl_xor_sync(0xffffffffu, d, 1));
}
__device__ __forceinline__ int column_of_lane(int lane) { return (lane >> 1) & 15; }   // h16*8 + h8*4 + h4*2 + h2
// Bit pattern of |a - b| for two non-negative doubles, formed with integer instructions (see
// final_value for why): the mantissas are aligned at bit 62, subtracted and renormalised; the
// result is within one ulp of the exact difference (truncated, never above it).  Inf/NaN -> 0,
// the reference's `abs(a - b) > eps` is False for NaN as well.
__device__ __forceinline__ unsigned long long absdiff_bits(double a, double b) {
  const unsigned long long A = (unsigned long long)__double_as_longlong(a), B = (unsigned long long)__double_as_longlong(b);
  const unsigned long long hi = A > B ? A : B, lo = A > B ? B : A;
  const int eh = (int)(hi >> 52), el = (int)(lo >> 52);
  const int sh = eh - el;
  const unsigned long long mh = ((hi & 0xfffffffffffffull) | (1ull << 52)) << 10;
  const unsigned long long ml = (el > 0 && sh < 63) ? (((lo & 0xfffffffffffffull) | (1ull << 52)) << 10) >> sh : 0ull;
  const unsigned long long d = mh - ml;
  const int z = __clzll((long long)d) & 63;
  const int be = eh + 1 - z;
  const unsigned long long bits = ((unsigned long long)(unsigned)(be - 1) << 52) + ((d << z) >> 11);
  return (eh == 0 || eh >= 0x7ff || d == 0ull || be <= 0) ? 0ull : bits;
}

struct Params {
  int layout;
  int64_t M, R, K;
  srk_rowbound in_rowbound;
  // MID
  uint8_t* out_planes; int64_t ld_outp; int64_t out_plane_stride;
  srk_rowbound out_rowbound;
  // FINAL
  const double* g_a; const double* g_v;
  const void* counts; int64_t ld_counts; int add_counts, use_evidence, counts32;
  int out_counts8;             // COUNTS: clip to 255 and store uint8 (the evidence counts of SimRank.py:315)
  double* out_f64; int64_t ld_out; int64_t diag_offset;
  double* mirror_out; int64_t ld_mirror; int64_t mirror_col0;
  uint32_t* rowmax_hi;
  uint32_t* mirror_rowmax_hi;
  EpilogueDev epi;
  double* maxdiff; double* maxoff;
  // COUNTS
  void* out_counts; int64_t ld_out_counts;
  // schedule
  int tiles_j, tiles_r, group_j, total_tiles;
  int kblock;
  // lockstep of the CTA pairs (see "lockstep" in the producer): units of `sync_kb` K-blocks
  unsigned int* sync; int sync_units_per_tile, sync_kb, sync_lag;
  int flags;                   // SRK_X2_FLAGS (A/B profiling): 1 = operand loads, 2 = epilogue streams with the default L2 policy
  unsigned long long* trace;   // SRK_X2_TRACE (diagnostics): [clusters][trace_tiles][4] globaltimer at mainloop start/end, epilogue start/end
  int trace_tiles;
};

// Walks the pair tiles in the order: bands of `group_j` row blocks; inside a band the column
// blocks are the outer loop and the row blocks the inner one, so the ~74 concurrently running
// pairs share a compact block of operand panels in L2.  In the symmetric layout only tiles that
// contain an element with j <= r are visited.
template <int RT>
struct TileWalk {
  int jb, rb, band_lo, band_hi, tiles_j, tiles_r, group_j;
  bool sym, done;
  __device__ __forceinline__ bool needed(int j, int r) const { return !sym || (j * 256 <= r * RT + RT - 1); }
  __device__ __forceinline__ void enter_band(int lo) {
    band_lo = lo;
    band_hi = min(lo + group_j, tiles_j);
    jb = lo;
    rb = sym ? (lo * 256) / RT : 0;
    if (lo >= tiles_j || rb >= tiles_r) done = true;
  }
  __device__ __forceinline__ void init(const Params& p) {
    tiles_j = p.tiles_j; tiles_r = p.tiles_r; group_j = p.group_j;
    sym = p.layout == SRK_X2_SYMMETRIC;
    done = false;
    enter_band(0);
  }
  __device__ __forceinline__ void step() {
    ++jb;
    if (jb < band_hi && needed(jb, rb)) return;
    ++rb;
    jb = band_lo;
    if (rb >= tiles_r) enter_band(band_hi);
  }
  __device__ __forceinline__ void advance(int n) {
    for (int i = 0; i < n && !done; ++i) step();
  }
};

template <int NS, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
i8x2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_v, const Params p) {
  using C = Cfg<NS>;
  constexpr int RT = C::RT, N = C::N, kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  double* colfac = reinterpret_cast<double*>(smem + kStages * C::kStageBytes);       // [2][2][RT]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::kStageBytes + C::kColfacBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  double* mirror_stage = reinterpret_cast<double*>(smem + kStages * C::kStageBytes + C::kColfacBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int kblocks = (int)((p.K + BK - 1) / BK);
  const int my_tiles = p.total_tiles > cluster_id ? (p.total_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_v);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                  // both CTAs' barriers and TMEM are ready
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      TileWalk<RT> tw;
      tw.init(p);
      tw.advance(cluster_id);
      int stage = 0; uint32_t phase = 0;
      const uint64_t pol = (p.flags & 1) ? policy_evict_normal() : policy_evict_last();
      // The A8 row panels of a band (group_j x 256 rows) are re-read by every wave of the band, the
      // V panels only by the pairs of one wave (which are at the same k: lockstep).  SRK_X2_FLAGS 4 / 8
      // (A/B profiling) take the V loads out of the evict_last class (normal / evict_first), so that the
      // band of A8 can stay resident in L2 across the waves.
      const uint64_t polv = (p.flags & 8) ? policy_evict_first() : ((p.flags & 4) ? policy_evict_normal() : pol);
      unsigned int peek = 0u;                              // lockstep: counter of unit peek_unit, read ahead
      int peek_unit = -1;
      for (int t = 0; t < my_tiles; ++t) {
        const int j0 = tw.jb * 256 + (int)cta * BMC;
        const int r0 = tw.rb * RT;
        int sync_next = 0, peek_next = p.sync_kb >> 1, unit_in_tile = 0;   // K-blocks of the next boundary / look-ahead
        for (int kb = 0; kb < kblocks; ++kb) {
          if (p.sync && cta == 0) {
            // Lockstep.  The ~74 pairs of a wave read the same few operand panels; they only find
            // each other's lines in L2 while they are at about the same k.  Left alone they drift
            // apart by several tiles (memory-bound pairs run at slightly different speeds and nothing
            // ever re-aligns them) and every pair streams its panels from DRAM: measured 140 GB per
            // FINAL launch instead of 35 GB.  So progress is counted in units of sync_kb K-blocks and a
            // pair starts unit u only after every pair that has a unit u-lag started it.  The counter
            // is read half a unit ahead (the load is in flight while the next K-blocks are issued), so
            // in step the boundary costs nothing; only a pair that runs ahead polls.  The wait is
            // bounded: this is a performance hint, never a correctness dependency.
            if (kb == sync_next) {
              const int unit = t * p.sync_units_per_tile + unit_in_tile;
              sync_next += p.sync_kb;
              red_add_gpu(p.sync + unit);
              const int prev = unit - p.sync_lag;
              if (prev >= 0) {
                const int left = p.total_tiles - (prev / p.sync_units_per_tile) * num_clusters;
                const unsigned int expect = (unsigned int)(left < num_clusters ? left : num_clusters);
                if (!(peek_unit == prev && peek >= expect) && ld_relaxed_gpu(p.sync + prev) < expect) {
                  const unsigned long long t_in = globaltimer_ns();
                  while (ld_relaxed_gpu(p.sync + prev) < expect && globaltimer_ns() - t_in < 1000000ull) {}
                }
              }
            } else if (kb == peek_next) {
              peek_next += p.sync_kb;
              peek_unit = t * p.sync_units_per_tile + unit_in_tile + 1 - p.sync_lag;
              ++unit_in_tile;
              if (peek_unit >= 0) peek = ld_relaxed_gpu(p.sync + peek_unit);
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * C::kStageBytes;
          uint8_t* sb = sa + BMC * BK;
          if (cta == 0) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
          tma_load_2d(sa, &map_a, &full_bar[stage], pol, kb * BK, j0);
#pragma unroll
          for (int b = 0; b < C::kBoxes; ++b) {
            const int nrow = (int)cta * C::NH + b * C::BR;         // row of the concatenated N operand
            const int plane = nrow / RT, rr = nrow % RT;
            if (p.kblock > 0) {
              const int k = kb * BK;
              tma_load_4d(sb + b * C::BR * BK, &map_v, &full_bar[stage], polv, k % p.kblock, r0 + rr, k / p.kblock, plane);
            } else {
              tma_load_3d(sb + b * C::BR * BK, &map_v, &full_bar[stage], polv, kb * BK, r0 + rr, plane);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tw.advance(num_clusters);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && cta == 0) {
      constexpr uint32_t idesc = make_idesc<N>();
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int b = t & 1;
        mbar_wait(&tmem_empty[b], ((uint32_t)(t >> 1) & 1u) ^ 1u);    // both CTAs drained this buffer
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(b * kAccStride);
        if (p.trace && t < p.trace_tiles) p.trace[((size_t)cluster_id * p.trace_tiles + t) * 4] = globaltimer_ns();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * C::kStageBytes);
          const uint64_t desc_a = make_desc(sa);
          const uint64_t desc_b = make_desc(sa + BMC * BK);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            mma_i8_pair(acc, desc_a + (uint64_t)((k * UMMA_K) >> 4), desc_b + (uint64_t)((k * UMMA_K) >> 4), idesc,
                        (kb | k) ? 1u : 0u);
          mma_commit_pair(&empty_bar[stage]);                 // slot reusable in both CTAs
          if (kb == kblocks - 1) mma_commit_pair(&tmem_full[b]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (p.trace && t < p.trace_tiles) p.trace[((size_t)cluster_id * p.trace_tiles + t) * 4 + 1] = globaltimer_ns();
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - kEpiWarp0;                     // == warp % 4: TMEM lane quarter
    const int et = ew * 32 + lane;                       // 0..127: TMEM lane = row of this CTA's half
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
    TileWalk<RT> tw;
    tw.init(p);
    tw.advance(cluster_id);
    unsigned long long dmax = 0ull, omax = 0ull;         // bit patterns of non-negative doubles
    const bool sym = p.layout == SRK_X2_SYMMETRIC;
    const bool trans = p.layout == SRK_X2_TRANSPOSED;
    const bool have_old = p.epi.s_old != nullptr;
    const uint64_t spol = (p.flags & 2) ? policy_evict_normal() : policy_evict_first();
    // everything the vectorised paths assume about the caller's buffers (uniform over the grid)
    bool fast_ok = false;
    if (MODE == SRK_X2_FINAL) {
      fast_ok = have_old && !p.epi.prior && !p.epi.evidence &&
                ((reinterpret_cast<uintptr_t>(p.out_f64) | reinterpret_cast<uintptr_t>(p.epi.s_old)) & 15) == 0 &&
                ((p.ld_out | p.epi.ld_s_old) & 1) == 0;
      if (p.counts) fast_ok = fast_ok && (reinterpret_cast<uintptr_t>(p.counts) & 15) == 0 && (p.ld_counts & (p.counts32 ? 3 : 7)) == 0;
      if (p.mirror_out)
        fast_ok = fast_ok && (reinterpret_cast<uintptr_t>(p.mirror_out) & 15) == 0 && ((p.ld_mirror | p.mirror_col0) & 1) == 0;
    }
    for (int t = 0; t < my_tiles; ++t) {
      const int b = t & 1;
      const int64_t jc0 = (int64_t)tw.jb * 256 + (int64_t)cta * BMC;        // first A8 row of this CTA's half
      const int64_t j = jc0 + et;                                           // row of A8 owned by this thread
      const int64_t r0 = (int64_t)tw.rb * RT;
      const bool jvalid = j < p.M;
      // per-column factors of this tile (double-buffered; one named barrier per tile)
      double* cf = colfac + (size_t)b * RT;                                 // [2][RT] doubles
      int* shv = reinterpret_cast<int*>(colfac + 2 * RT) + (size_t)b * RT;  // [2][RT] ints
      const unsigned long long* mcf = reinterpret_cast<const unsigned long long*>(cf);   // FINAL: mantissas
      if (MODE != SRK_X2_COUNTS) {
        for (int c = et; c < RT; c += 128) {
          const int64_t r = r0 + c;
          double f = 0.0;
          int sh = 0;
          if (r < p.R) {
            const double bin = row_bound(p.in_rowbound, r);
            if (MODE == SRK_X2_MID) {
              f = bin > 0.0 ? bin : 0.0;
            } else {
              const int fr = pow2_exponent<NS>(bin);
              int s = fr == kNoBound ? 0 : 8 * NS - fr, dl = 0;
              if (s < 0) { dl = -s > 17 ? 17 : -s; s = 0; }
              unsigned long long m_cf;
              int e_cf;
              split_f64(p.epi.coef * p.g_v[r] * pow2(-s), m_cf, e_cf);
              sh = pack_shift(s, dl, e_cf);
              f = __longlong_as_double((long long)m_cf);        // FINAL keeps the mantissa bits in the slot
            }
          }
          cf[c] = f;
          shv[c] = sh;
        }
        epi_bar_sync();
      }
      unsigned long long rmax = 0ull;                     // FINAL, row-major layouts: max of row j inside this tile
      unsigned long long m_gj = 0ull;                     // FINAL: g_a[j] = m_gj / 2^63 * 2^e_gj
      int e_gj = 0;
      int fj = kNoBound;                                  // MID: exponent of the power-of-two bound of row j
      if (jvalid) {
        if (MODE == SRK_X2_MID) fj = pow2_exponent<NS>(row_bound(p.out_rowbound, j));
        if (MODE == SRK_X2_FINAL) split_f64(p.g_a[j], m_gj, e_gj);
      }
      if (MODE == SRK_X2_FINAL && !trans && jvalid && have_old && !(sym && j > r0 + RT - 1)) {
        // The accumulator of this tile is still being computed: pull the rows of S_old and of the
        // counts that its epilogue will read into L2 now, so that the eight dependent
        // load -> compute -> store rounds below see L2 latency instead of DRAM latency.  (An
        // epilogue that outlasts a mainloop stalls the MMA issuer, the CTA pairs drift apart in k
        // and stop sharing operand panels in L2 -- measured as 2x the DRAM traffic of MID.)
        const int64_t cols = min((int64_t)RT, p.R - r0);
        const char* sp = reinterpret_cast<const char*>(p.epi.s_old + j * p.epi.ld_s_old + r0);
        for (int64_t o = 0; o < cols * 8; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + o));
        if (p.counts) {
          const int cb = p.counts32 ? 4 : 2;
          const char* cp = reinterpret_cast<const char*>(p.counts) + (j * p.ld_counts + r0) * cb;
          for (int64_t o = 0; o < cols * cb; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + o));
        }
      }
      mbar_wait(&tmem_full[b], (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const uint32_t acc = lane_base + (uint32_t)(b * kAccStride);
      const bool tracing = p.trace && t < p.trace_tiles && cta == 0 && et == 0;
      if (tracing) p.trace[((size_t)cluster_id * p.trace_tiles + t) * 4 + 2] = globaltimer_ns();

      // MID keeps the packed planes of its whole row (RT bytes per plane) in registers until the tile is
      // done, then the warp writes them out as whole lines (below); that needs a fully unrolled loop.
      uint4 wq[MODE == SRK_X2_MID ? NS : 1][RT / 16];
      if (MODE == SRK_X2_MID) {
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int c = 0; c < RT / 16; ++c) wq[s][c] = make_uint4(0u, 0u, 0u, 0u);
      }
      constexpr int kChunkUnroll = MODE == SRK_X2_MID ? RT / 16 : 1;
#pragma unroll kChunkUnroll
      for (int c0 = 0; c0 < RT; c0 += 16) {
        __syncwarp();                                     // reconverge after the divergent tails below
        uint32_t a[NS][16];
#pragma unroll
        for (int s = 0; s < NS; ++s) tmem_ld16(acc + s * RT + c0, a[s]);
        tmem_wait_ld();
        if (c0 + 16 >= RT) {                              // last read of this buffer: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[b], 0);
        }
        const int64_t rc = r0 + c0;                       // first V row of this chunk
        if (rc >= p.R) continue;                          // warp-uniform

        if (MODE == SRK_X2_COUNTS) {
          if (!jvalid) continue;
          if (p.out_counts8) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int x = 0; x < 16; ++x)
              w[x >> 2] |= ((rc + x < p.R) ? min(a[0][x], 255u) : 0u) << (8 * (x & 3));
            uint8_t* o = reinterpret_cast<uint8_t*>(p.out_counts) + j * p.ld_out_counts + rc;
            if (rc + 16 <= p.ld_out_counts && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
              *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
#pragma unroll
              for (int x = 0; x < 16; ++x)
                if (rc + x < p.R) o[x] = (uint8_t)((w[x >> 2] >> (8 * (x & 3))) & 0xffu);
            }
            continue;
          }
          if (p.counts32) {
            uint32_t* o = reinterpret_cast<uint32_t*>(p.out_counts) + j * p.ld_out_counts + rc;
            if (rc + 16 <= p.ld_out_counts && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
              for (int x = 0; x < 16; x += 4)
                reinterpret_cast<uint4*>(o)[x >> 2] = make_uint4(rc + x < p.R ? a[0][x] : 0u, rc + x + 1 < p.R ? a[0][x + 1] : 0u,
                                                                 rc + x + 2 < p.R ? a[0][x + 2] : 0u, rc + x + 3 < p.R ? a[0][x + 3] : 0u);
            } else {
#pragma unroll
              for (int x = 0; x < 16; ++x)
                if (rc + x < p.R) o[x] = a[0][x];
            }
            continue;
          }
          uint32_t w[8];
#pragma unroll
          for (int x = 0; x < 16; x += 2) {
            const uint32_t lo = (rc + x < p.R) ? min(a[0][x], 65535u) : 0u;
            const uint32_t hi = (rc + x + 1 < p.R) ? min(a[0][x + 1], 65535u) : 0u;
            w[x >> 1] = lo | (hi << 16);
          }
          uint16_t* o = reinterpret_cast<uint16_t*>(p.out_counts) + j * p.ld_out_counts + rc;
          if (rc + 16 <= p.ld_out_counts && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
          } else {
#pragma unroll
            for (int x = 0; x < 16; ++x)
              if (rc + x < p.R) o[x] = (uint16_t)((w[x >> 1] >> (16 * (x & 1))) & 0xffffu);
          }
          continue;
        }

        if (MODE == SRK_X2_MID) {
          uint32_t w[NS][4];
#pragma unroll
          for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int x = 0; x < 4; ++x) w[s][x] = 0u;
          if (jvalid && fj != kNoBound) {
#pragma unroll
            for (int x = 0; x < 16; ++x) {
              const uint32_t q = mid_quant<NS>(combine<NS>(a, x), cf[c0 + x], fj);
#pragma unroll
              for (int s = 0; s < NS; ++s) w[s][x >> 2] |= ((q >> (8 * (NS - 1 - s))) & 0xffu) << (8 * (x & 3));
            }
          }
#pragma unroll
          for (int s = 0; s < NS; ++s) wq[MODE == SRK_X2_MID ? s : 0][c0 / 16] = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
          continue;
        }

        // ---------------------------------------------------------------- FINAL
        // keys of the values this lane mirrors into rows rc + x (symmetric layout); their column
        // maxima are taken by ALL lanes of the warp after the possibly divergent paths below
        uint32_t key[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) key[x] = 0u;
        bool staged = false;                              // this lane's mirror values are in the warp's staging tile
        const int64_t jd = j - p.diag_offset;             // V row that sits on the diagonal with j
        // Chunks whose 16 elements are all in range take the vectorised paths.  In the row-major
        // layouts that includes the chunk that holds the diagonal element of row j: its loads are the
        // same 128-bit loads, only the stores are predicated (symmetric layout: elements left of the
        // diagonal belong to the mirror image of another tile).  A tile that crosses the diagonal so
        // costs about as much as any other -- with the pairs in lockstep one slow epilogue per wave
        // would otherwise set the pace of every wave.
        const bool in_range = fast_ok && rc + 16 <= p.R;
        const bool fast_t = in_range && trans && (jd < rc || jd > rc + 15);
        const bool fast_n = in_range && !trans && !(sym && jd > rc + 15);
        double v[16];
        if (!jvalid) {
          // rows past the end of A8: nothing to store, but stay for the warp-wide reduction
        } else if (fast_n) {
          const int xd = jd < rc ? -1 : (jd > rc + 15 ? 16 : (int)(jd - rc));   // diagonal position, -1 / 16 = none
          const int x_lo = sym ? xd : -1;                 // symmetric layout: elements x >= xd are this tile's
          uint32_t cv[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) cv[x] = 0u;
          if (p.counts && p.counts32) {
            const uint4* cp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(p.counts) + j * p.ld_counts + rc);
#pragma unroll
            for (int x = 0; x < 4; ++x) { const uint4 t4 = ld_stream_u32x4(cp + x, spol); cv[4 * x] = t4.x; cv[4 * x + 1] = t4.y; cv[4 * x + 2] = t4.z; cv[4 * x + 3] = t4.w; }
          } else if (p.counts) {
            const uint4* cp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.counts) + j * p.ld_counts + rc);
            const uint4 t0 = ld_stream_u32x4(cp, spol), t1 = ld_stream_u32x4(cp + 1, spol);
            const uint32_t cw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
            for (int x = 0; x < 16; ++x) cv[x] = (cw[x >> 1] >> (16 * (x & 1))) & 0xffffu;
          }
          double so[16];
          const double2* sp = reinterpret_cast<const double2*>(p.epi.s_old + j * p.epi.ld_s_old + rc);
#pragma unroll
          for (int x = 0; x < 8; ++x) { const double2 d2 = ld_stream_f64x2(sp + x, spol); so[2 * x] = d2.x; so[2 * x + 1] = d2.y; }
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const uint32_t cnt = cv[x];
            double val = final_value(combine<NS>(a, x), p.add_counts ? cnt : 0u, shv[c0 + x], mcf[c0 + x], m_gj, e_gj);
            if (p.use_evidence) val *= evidence_factor(cnt);
            if (x == xd) val = 1.0;
            else if (x > x_lo) rmax = umax64(rmax, (unsigned long long)__double_as_longlong(val));
            if (x >= x_lo) dmax = umax64(dmax, absdiff_bits(val, so[x]));
            v[x] = val;
          }
          if (x_lo <= 0) {                                // every element of the chunk is stored
            double2* op = reinterpret_cast<double2*>(p.out_f64 + j * p.ld_out + rc);
#pragma unroll
            for (int x = 0; x < 8; ++x) st_stream_f64x2(op + x, v[2 * x], v[2 * x + 1], spol);
          } else {
            double* op = p.out_f64 + j * p.ld_out + rc;
#pragma unroll
            for (int x = 0; x < 16; ++x)
              if (x >= x_lo) st_stream_f64(op + x, v[x], spol);
          }
          if (sym) {
            double* mp = p.out_f64 + rc * p.ld_out + j;   // mirror: 32 lanes write 256 contiguous bytes per x
#pragma unroll
            for (int x = 0; x < 16; ++x)
              if (x > x_lo) { st_stream_f64(mp + x * p.ld_out, v[x], spol); key[x] = hi_key(v[x]); }
          }
        } else if (fast_t) {
          // element (row rc + x, column j): every access is coalesced across the lanes
          uint32_t cnt[16];
          double so[16];
          const double* sp = p.epi.s_old + rc * p.epi.ld_s_old + j;
#pragma unroll
          for (int x = 0; x < 16; ++x) so[x] = ld_stream_f64(sp + x * p.epi.ld_s_old, spol);
          if (p.counts) {
#pragma unroll
            for (int x = 0; x < 16; ++x) cnt[x] = load_count(p.counts, (rc + x) * p.ld_counts + j, p.counts32);
          } else {
#pragma unroll
            for (int x = 0; x < 16; ++x) cnt[x] = 0u;
          }
          double* op = p.out_f64 + rc * p.ld_out + j;
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            double val = final_value(combine<NS>(a, x), p.add_counts ? cnt[x] : 0u, shv[c0 + x], mcf[c0 + x], m_gj, e_gj);
            if (p.use_evidence) val *= evidence_factor(cnt[x]);
            rmax = umax64(rmax, (unsigned long long)__double_as_longlong(val));
            dmax = umax64(dmax, absdiff_bits(val, so[x]));
            st_stream_f64(op + x * p.ld_out, val, spol);
            key[x] = hi_key(val);
            v[x] = val;
          }
          if (p.mirror_out) {
            // (row j, columns rc..rc+15) of the GPU that owns row j: 128 contiguous bytes per lane.
            // Stored lane by lane that is 32 x 16 B pieces per instruction; over NVLink every piece is
            // its own write packet.  The warp stages the 32 x 16 tile in shared memory instead and
            // writes it out as whole 128-byte lines, four rows per instruction (below).
            double2* sm = reinterpret_cast<double2*>(mirror_stage + (ew * 32 + lane) * C::kMirrorRow);
#pragma unroll
            for (int x = 0; x < 8; ++x) sm[x] = make_double2(v[2 * x], v[2 * x + 1]);
            staged = true;
          }
        } else if (!(sym && jd > rc + 15)) {              // (strictly below the diagonal: mirrored from above)
          // ---- general path: chunks that touch the diagonal or an edge, priors, uint8 evidence
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const int64_t r = rc + x;
            const bool live = r < p.R && !(sym && jd > r);
            if (!live) continue;
            const int64_t idx_o = trans ? r * p.ld_out + j : j * p.ld_out + r;
            uint32_t cnt = 0u;
            if (p.counts) cnt = load_count(p.counts, trans ? r * p.ld_counts + j : j * p.ld_counts + r, p.counts32);
            double val = final_value(combine<NS>(a, x), p.add_counts ? cnt : 0u, shv[c0 + x], mcf[c0 + x], m_gj, e_gj);
            if (p.use_evidence) val *= evidence_factor(cnt);
            else if (p.epi.evidence)
              val *= evidence_factor(trans ? p.epi.evidence[r * p.epi.ld_evidence + j] : p.epi.evidence[j * p.epi.ld_evidence + r]);
            if (p.epi.prior)
              val = (1.0 - p.epi.lambda) * val +
                    p.epi.lambda * (trans ? p.epi.prior[r * p.epi.ld_prior + j] : p.epi.prior[j * p.epi.ld_prior + r]);
            if (r == jd) val = 1.0;
            else if (val > 0.0) {
              rmax = umax64(rmax, (unsigned long long)__double_as_longlong(val));
              if (sym || trans) key[x] = hi_key(val);     // symmetric: r > jd here, the value is mirrored to row r
            }
            if (have_old) {
              const double so = trans ? p.epi.s_old[r * p.epi.ld_s_old + j] : p.epi.s_old[j * p.epi.ld_s_old + r];
              const double d = fabs(val - so);
              if (d > 0.0) dmax = umax64(dmax, (unsigned long long)__double_as_longlong(d));   // NaN compares false
            }
            p.out_f64[idx_o] = val;
            if (sym && jd < r) p.out_f64[r * p.ld_out + j] = val;
            if (trans && p.mirror_out && r != jd) p.mirror_out[j * p.ld_mirror + p.mirror_col0 + r] = val;
          }
        }
        if ((sym || trans) && p.rowmax_hi) {              // warp-uniform
          // symmetric: keys of the values mirrored into rows rc + x; transposed: rows rc + x of the output
          __syncwarp();
          const uint32_t cm = column_max16(key, lane);
          if (!(lane & 1) && cm > 1u) atomicMax(p.rowmax_hi + rc + column_of_lane(lane), cm);
        }
        if (trans && p.mirror_out) {                      // warp-uniform
          const unsigned smask = __ballot_sync(0xffffffffu, staged);
          if (smask) {
            const int64_t jw = jc0 + ew * 32;             // A8 row of lane 0
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = 4 * i + (lane >> 3), piece = lane & 7;
              if ((smask >> row) & 1u) {
                const double2 d = *reinterpret_cast<const double2*>(mirror_stage + (ew * 32 + row) * C::kMirrorRow + 2 * piece);
                st_stream_f64x2(reinterpret_cast<double2*>(p.mirror_out + (jw + row) * p.ld_mirror + p.mirror_col0 + rc) + piece,
                                d.x, d.y, spol);
              }
            }
            __syncwarp();                                 // the tile is rewritten by the next chunk
          }
        }
      }
      if (MODE == SRK_X2_MID) {
        // Row j of the tile is RT contiguous bytes per plane, but held by ONE lane: stored from there
        // every instruction scatters 32 x 16 B over 32 lines (and, when the row panel of U lives on
        // another GPU, over 32 NVLink write packets: 2 ms per iteration at 2 GPUs).  Each plane goes
        // through the warp's staging tile and leaves as whole rows, 32 / (RT / 16) rows per instruction.
        constexpr int kPieces = RT / 16, kRows = 32 / kPieces;
        uint8_t* stg = reinterpret_cast<uint8_t*>(mirror_stage) + ew * (32 * C::kMirrorRow * 8);
        const int64_t jw = jc0 + ew * 32;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          __syncwarp();
#pragma unroll
          for (int c = 0; c < kPieces; ++c)
            *reinterpret_cast<uint4*>(stg + lane * (C::kMirrorRow * 8) + c * 16) = wq[MODE == SRK_X2_MID ? s : 0][c];
          __syncwarp();
#pragma unroll
          for (int i = 0; i < kPieces; ++i) {
            const int row = i * kRows + lane / kPieces, piece = lane % kPieces;
            const int64_t jj = jw + row, rcol = r0 + piece * 16;
            if (jj < p.M && rcol < p.R)
              st_stream_u32x4(reinterpret_cast<uint4*>(p.out_planes + s * p.out_plane_stride + jj * p.ld_outp + rcol),
                              *reinterpret_cast<const uint4*>(stg + row * (C::kMirrorRow * 8) + piece * 16), spol);
          }
        }
      }
      if (tracing) p.trace[((size_t)cluster_id * p.trace_tiles + t) * 4 + 3] = globaltimer_ns();
      if (MODE == SRK_X2_FINAL && rmax) {
        omax = umax64(omax, rmax);
        // row j of the result (row-major layouts) / of the mirrored block (transposed layout)
        uint32_t* rk = trans ? (p.mirror_out ? p.mirror_rowmax_hi : nullptr) : p.rowmax_hi;
        if (rk) atomicMax(rk + j, (uint32_t)(rmax >> 32) + 1u);
      }
      tw.advance(num_clusters);
    }
    if (MODE == SRK_X2_FINAL) {
      double dm = __longlong_as_double((long long)dmax), om = __longlong_as_double((long long)omax);
      dm = warp_max(dm);
      om = warp_max(om);
      if (lane == 0) {
        if (p.maxdiff && dm > 0.0) atomic_max_nonneg(p.maxdiff, dm);
        if (p.maxoff && om > 0.0) atomic_max_nonneg(p.maxoff, om);
      }
    }
  }

  // teardown: no CTA may exit (or free TMEM) while its peer can still signal it
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(SRK_ERR_CUDA, "%s", "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SRK_ERR_CUDA, "cuTensorMapEncodeTiled failed%s (code %lld)", "", (long long)r);
  return SRK_OK;
}

static int count_tiles(int tiles_j, int tiles_r, int rt, bool sym) {
  if (!sym) return tiles_j * tiles_r;
  long long total = 0;
  for (int jb = 0; jb < tiles_j; ++jb) {
    const int first = (jb * 256) / rt;
    if (first < tiles_r) total += tiles_r - first;
  }
  return (int)total;
}

template <int NS, int MODE>
static int launch(const srk_x2_args& a, cudaStream_t st) {
  using C = Cfg<NS>;
  CUtensorMap map_a, map_v;
  int rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.M};
    cuuint64_t strides[1] = {(cuuint64_t)a.lda};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BMC};
    rc = make_map(&map_a, a.A8, 2, dims, strides, box);
  }
  if (rc) return rc;
  if (a.in_kblock > 0) {
    cuuint64_t dims[4] = {(cuuint64_t)a.in_kblock, (cuuint64_t)a.R, (cuuint64_t)(a.K / a.in_kblock), (cuuint64_t)NS};
    cuuint64_t strides[3] = {(cuuint64_t)a.ld_in, (cuuint64_t)a.in_kblock_stride, (cuuint64_t)a.in_plane_stride};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)C::BR, 1, 1};
    rc = make_map(&map_v, a.in_planes, 4, dims, strides, box);
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)a.K, (cuuint64_t)a.R, (cuuint64_t)NS};
    const int64_t whole = a.ld_in * a.R;                      // a 1-plane operand still needs a legal stride
    cuuint64_t strides[2] = {(cuuint64_t)a.ld_in, (cuuint64_t)(NS > 1 ? a.in_plane_stride : whole)};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)C::BR, 1};
    rc = make_map(&map_v, a.in_planes, 3, dims, strides, box);
  }
  if (rc) return rc;
  Params p;
  memset(&p, 0, sizeof(p));
  p.layout = MODE == SRK_X2_FINAL ? a.layout : SRK_X2_DIRECT;
  p.M = a.M; p.R = a.R; p.K = a.K;
  p.in_rowbound = a.in_rowbound;
  p.out_planes = a.out_planes; p.ld_outp = a.ld_outp; p.out_plane_stride = a.out_plane_stride;
  p.out_rowbound = a.out_rowbound;
  p.g_a = a.g_a; p.g_v = a.g_v;
  p.counts = a.counts; p.ld_counts = a.ld_counts; p.add_counts = a.add_counts; p.use_evidence = a.use_evidence;
  p.counts32 = a.counts_bits == 32;
  p.out_counts8 = MODE == SRK_X2_COUNTS && a.counts_bits == 8;
  p.out_f64 = a.out_f64; p.ld_out = a.ld_out; p.diag_offset = a.diag_offset;
  p.mirror_out = a.mirror_out; p.ld_mirror = a.ld_mirror; p.mirror_col0 = a.mirror_col0;
  p.rowmax_hi = a.rowmax_hi;
  p.mirror_rowmax_hi = a.mirror_rowmax_hi;
  p.epi = to_dev(a.epi);
  p.maxdiff = a.epi.maxdiff; p.maxoff = a.epi.maxoff;
  p.out_counts = a.out_counts; p.ld_out_counts = a.ld_out_counts;
  p.tiles_j = (int)((a.M + 255) / 256);
  p.tiles_r = (int)((a.R + C::RT - 1) / C::RT);
  p.group_j = 8;
  if (const char* e = getenv("SRK_X2_GROUP")) { const int gj = atoi(e); if (gj >= 1 && gj <= 64) p.group_j = gj; }   // A/B profiling
  p.total_tiles = count_tiles(p.tiles_j, p.tiles_r, C::RT, p.layout == SRK_X2_SYMMETRIC);
  p.kblock = (int)a.in_kblock;
  { const char* e = getenv("SRK_X2_FLAGS"); p.flags = e ? atoi(e) : 0; }

  auto kern = i8x2_kernel<NS, MODE>;
  SRK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  int dev = 0, sms = 0;
  SRK_CUDA_OK(cudaGetDevice(&dev));
  SRK_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(sms / 2 * 2), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  SRK_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  if (max_clusters < 1) return fail(SRK_ERR_CUDA, "%s", "no CTA pair can be resident on this device");
  int clusters = sms / 2 < max_clusters ? sms / 2 : max_clusters;
  if (p.total_tiles < clusters) clusters = p.total_tiles;
  if (clusters < 1) return SRK_OK;
  cfg.gridDim = dim3((unsigned)(2 * clusters), 1, 1);
  {
    // lockstep workspace: one counter per (tile index, eighth of the K loop)
    const int kblocks = (int)((a.K + BK - 1) / BK);
    const char* env = getenv("SRK_X2_LOCKSTEP");                       // "0": let the pairs run free (A/B profiling)
    const bool off = env && env[0] == '0';
    int per_tile = kblocks >= 256 ? 8 : (kblocks >= 32 ? kblocks / 32 : 1), lag = 1;   // units of >= 32 K-blocks
    if (env && env[0] != '0' && kblocks >= 64) sscanf(env, "%d,%d", &per_tile, &lag);   // A/B profiling: "units,lag"
    if (per_tile < 1 || per_tile > kblocks / 2) per_tile = 1;          // a unit spans >= 2 K-blocks (boundary + look-ahead)
    if (lag < 1) lag = 1;
    const int64_t tiles_per_cluster = (p.total_tiles + clusters - 1) / clusters;
    const int64_t need = tiles_per_cluster * per_tile * 4;

    if (a.sync_ws && a.sync_ws_bytes >= need && clusters > 1 && kblocks >= 16 && !off) {
      p.sync = reinterpret_cast<unsigned int*>(a.sync_ws);
      p.sync_units_per_tile = per_tile;
      p.sync_lag = lag;
      p.sync_kb = (kblocks + per_tile - 1) / per_tile;
      SRK_CUDA_OK(cudaMemsetAsync(a.sync_ws, 0, (size_t)need, st));
    }
  }
  if (const char* tp = getenv("SRK_X2_TRACE")) {
    // diagnostics only: per-tile mainloop start/end times of every CTA pair, dumped after a blocking launch
    static int launch_no = 0;
    p.trace_tiles = (p.total_tiles + clusters - 1) / clusters;
    const size_t words = (size_t)clusters * p.trace_tiles * 4;
    SRK_CUDA_OK(cudaMalloc(&p.trace, words * 8));
    SRK_CUDA_OK(cudaMemsetAsync(p.trace, 0, words * 8, st));
    SRK_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, map_a, map_v, p));
    SRK_CUDA_OK(cudaStreamSynchronize(st));
    unsigned long long* h = (unsigned long long*)malloc(words * 8);
    SRK_CUDA_OK(cudaMemcpy(h, p.trace, words * 8, cudaMemcpyDeviceToHost));
    char name[512];
    snprintf(name, sizeof(name), "%s.%03d.m%d.ns%d.c%d.t%d.bin", tp, launch_no++, MODE, NS, clusters, p.trace_tiles);
    if (FILE* f = fopen(name, "wb")) { fwrite(h, 8, words, f); fclose(f); }
    free(h);
    cudaFree(p.trace);
    return SRK_OK;
  }
  SRK_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, map_a, map_v, p));
  return SRK_OK;
}

}  // namespace x2
}  // namespace srk

using namespace srk;

extern "C" int srk_i8_supported(void) {
  int cc = srk_device_cc();
  return cc >= 100 && cc < 103 ? 1 : 0;     // kind::i8 exists on sm_100a/sm_101a only
}

extern "C" int srk_x2_half(const srk_x2_args* a, void* stream) {
  SRK_REQUIRE(a, "null args");
  SRK_REQUIRE(a->in_planes && a->A8, "null operand");
  SRK_REQUIRE(a->M > 0 && a->R > 0 && a->K > 0, "empty problem");
  SRK_REQUIRE(a->ld_in % 16 == 0 && a->lda % 16 == 0, "operand strides must be multiples of 16");
  SRK_REQUIRE(((uintptr_t)a->in_planes % 16) == 0 && ((uintptr_t)a->A8 % 16) == 0, "operands must be 16-byte aligned");
  SRK_REQUIRE(a->lda >= a->K, "lda smaller than K");
  SRK_REQUIRE(a->M < (1ll << 31) - 512 && a->R < (1ll << 31) - 512, "too many rows for 32-bit tile coordinates");
  if (a->in_kblock > 0) {
    SRK_REQUIRE(a->in_kblock % 128 == 0 && a->K % a->in_kblock == 0 && a->ld_in >= a->in_kblock &&
                    a->in_kblock_stride % 16 == 0, "K-blocked operand: in_kblock must be a multiple of 128 dividing K");
  } else {
    SRK_REQUIRE(a->ld_in >= a->K, "ld_in smaller than K");
  }
  SRK_REQUIRE(a->K < (1ll << 22), "K too large for exact int32 accumulation");
  if (!srk_i8_supported()) return fail(SRK_ERR_UNSUPPORTED, "%s", "tcgen05 kind::i8 needs an sm_100 device");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->mode == SRK_X2_COUNTS) {
    SRK_REQUIRE(a->ns == 1 && a->out_counts && a->ld_out_counts >= a->R, "COUNTS needs ns=1 and a count output");
    SRK_REQUIRE(a->counts_bits == 0 || a->counts_bits == 8 || a->counts_bits == 16 || a->counts_bits == 32,
                "counts_bits must be 8, 16 or 32");
    return x2::launch<1, SRK_X2_COUNTS>(*a, st);
  }
  SRK_REQUIRE(a->ns == 1 || a->in_plane_stride % 16 == 0, "plane stride must be a multiple of 16");
  if (a->mode == SRK_X2_MID) {
    SRK_REQUIRE(a->out_planes, "MID needs output planes");
    SRK_REQUIRE(a->ld_outp % 16 == 0 && a->out_plane_stride % 16 == 0 && ((uintptr_t)a->out_planes % 16) == 0,
                "output planes must be 16-byte aligned with ld multiple of 16");
    SRK_REQUIRE(a->ld_outp >= a->R, "ld_outp smaller than R");
    switch (a->ns) {
      case 2: return x2::launch<2, SRK_X2_MID>(*a, st);
      case 3: return x2::launch<3, SRK_X2_MID>(*a, st);
      case 4: return x2::launch<4, SRK_X2_MID>(*a, st);
    }
    return fail(SRK_ERR_INVALID, "invalid argument: %s", "ns must be 2, 3 or 4");
  }
  if (a->mode == SRK_X2_FINAL) {
    SRK_REQUIRE(a->out_f64 && a->g_a && a->g_v, "FINAL needs out_f64, g_a, g_v");
    SRK_REQUIRE(a->layout == SRK_X2_DIRECT || a->layout == SRK_X2_SYMMETRIC || a->layout == SRK_X2_TRANSPOSED,
                "unknown layout");
    SRK_REQUIRE(!(a->use_evidence && a->epi.evidence), "evidence given twice (counts and epi.evidence)");
    SRK_REQUIRE(!(a->add_counts || a->use_evidence) || a->counts, "counts missing");
    SRK_REQUIRE(a->counts_bits == 0 || a->counts_bits == 16 || a->counts_bits == 32, "counts_bits must be 16 or 32");
    if (a->layout == SRK_X2_TRANSPOSED) {
      SRK_REQUIRE(a->ld_out >= a->M, "ld_out smaller than M (transposed layout)");
    } else {
      SRK_REQUIRE(a->ld_out >= a->R, "ld_out smaller than R");
    }
    SRK_REQUIRE(!a->mirror_rowmax_hi || a->mirror_out, "mirror_rowmax_hi without mirror_out");
    SRK_REQUIRE(!a->mirror_out || (a->layout == SRK_X2_TRANSPOSED && a->ld_mirror >= a->mirror_col0 + a->R),
                "mirror_out needs the transposed layout and ld_mirror >= mirror_col0 + R");
    if (a->layout == SRK_X2_SYMMETRIC)
      SRK_REQUIRE(a->M == a->R && a->diag_offset == 0 && a->epi.prior == nullptr,
                  "symmetric layout needs a square product without a prior");
    switch (a->ns) {
      case 2: return x2::launch<2, SRK_X2_FINAL>(*a, st);
      case 3: return x2::launch<3, SRK_X2_FINAL>(*a, st);
      case 4: return x2::launch<4, SRK_X2_FINAL>(*a, st);
    }
    return fail(SRK_ERR_INVALID, "invalid argument: %s", "ns must be 2, 3 or 4");
  }
  return fail(SRK_ERR_INVALID, "invalid argument: %s", "unknown mode");
}
