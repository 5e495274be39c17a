// Tensor-core half-product for the dense SimRank chain on sm_100a.
//
//   D[r, j] = sum_k V[r, k] * A8[j, k]
//
// V is a non-negative matrix held as NS uint8 fixed-point planes (V ~= q * rowbound / 256^NS),
// A8 is the 0/1 adjacency pattern.  Every plane is multiplied against A8 with
// tcgen05.mma.kind::i8 (u8 x u8 -> s32): the products and the K-long sums are EXACT integers, so
// the only rounding in a half-product is the re-quantisation of its result -- this is how the
// float64 reference (numpy dgemm, SimRank.py:139) is matched to ~1e-8 on a pipe that has no f64
// MMA.  See DESIGN.md "K1/K2".
//
// Structure (one CTA per SM, persistent over output tiles of BM x BN):
//   warp 0   TMA producer: one 3-D box (BK bytes x BM rows x NS planes) of V and one 2-D box
//            (BK x BN) of A8 per K-block into a kStages-deep 128B-swizzled smem ring
//   warp 1   MMA issuer: NS x (BK/32) tcgen05.mma per K-block into NS TMEM accumulators
//            (BM lanes x BN columns of s32 each); tcgen05.commit frees the smem slot
//   warp 2   TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld -> exact 64-bit recombination of the planes -> fused epilogue
//            MID   : + unit diagonal term, re-quantise, TRANSPOSED store through smem
//            FINAL : g_row*g_col, C, evidence, prior, diag<-1, max|dS|, f64 store (+ planes)
//            COUNTS: min(D,255) as uint8 (evidence counts A A^T)
#include <cuda.h>

#include "common.cuh"

namespace srk {
namespace i8 {

constexpr int BM = 128;        // rows of D per tile (= TMEM lanes)
constexpr int BK = 128;        // bytes of K per pipeline stage (= one 128B swizzle atom)
constexpr int UMMA_K = 32;     // K per tcgen05.mma for 8-bit operands
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;   // first epilogue warp (warp % 4 selects the TMEM lane quarter)
constexpr int kSmemLimit = 232448;

template <int NS, int BN>
struct Cfg {
  static constexpr int kStageBytes = NS * BM * BK + BN * BK;
  static constexpr int kStaging = 4 * NS * 32 * 32;                 // MID transpose buffers
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - 256 - kStaging) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kStaging + 256;
  static constexpr int kTmemCols = NS * BN <= 32 ? 32 : NS * BN <= 64 ? 64 : NS * BN <= 128 ? 128
                                   : NS * BN <= 256 ? 256 : 512;
  static_assert(NS * BN <= 512, "accumulators exceed TMEM");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
  static_assert(kStages >= 2, "not enough shared memory for a pipeline");
};

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, rows of 128 B, 128B swizzle: 8-row groups are 1024 B apart (SBO),
// LBO is unused for swizzled K-major layouts (1 as CUTLASS encodes it), descriptor version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::i8 instruction descriptor: D=s32, A=B=u8, both K-major, M=128, N=BN
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// exact value of sum_s acc_s * 256^(NS-1-s) as a double (< 2^53 for every supported shape)
template <int NS>
__device__ __forceinline__ double combine(const uint32_t (&a)[NS][16], int x) {
  long long v = 0;
#pragma unroll
  for (int s = 0; s < NS; ++s) v = (v << 8) + (long long)(int)a[s][x];
  return (double)v;
}

struct Params {
  int mode, unit_diag;
  int64_t R, N, K;
  srk_rowbound in_rowbound;
  const uint8_t* A8; int64_t lda;
  int64_t diag_offset;
  const double* g_row; const double* g_col;
  double* out_f64; int64_t ld_out;
  uint8_t* out_planes; int64_t ld_outp; int64_t out_plane_stride;
  srk_rowbound out_rowbound;
  EpilogueDev epi;
  double* maxdiff; double* maxoff;
  int tiles_m, tiles_n, group_m;
  int kblock;                  // > 0: V's K axis is split in blocks of `kblock` columns (4-D tensor map)
};

__device__ __forceinline__ void tile_coords(const Params& p, int tile, int& mb, int& nb) {
  // bands of `group_m` row-blocks, column-blocks fastest inside a band: the CTAs running at the
  // same time share a compact (group_m x ~148/group_m) block of operand panels in L2.
  const int band_tiles = p.group_m * p.tiles_n;
  const int band = tile / band_tiles;
  const int first_m = band * p.group_m;
  const int gm = min(p.group_m, p.tiles_m - first_m);
  const int in_band = tile - band * band_tiles;
  mb = first_m + in_band % gm;
  nb = in_band / gm;
}

template <int NS, int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
i8_half_kernel(const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_a,
               const Params p) {
  using C = Cfg<NS, BN>;
  constexpr int kStages = C::kStages;
  constexpr double kQ = (double)(1ull << (8 * NS));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kStages * C::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + C::kStaging);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int kblocks = (int)((p.K + BK - 1) / BK);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_v);
    tma_prefetch_desc(&map_a);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mb, nb;
        tile_coords(p, tile, mb, nb);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sv = stage_base + stage * C::kStageBytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          if (p.kblock > 0) {
            const int k = kb * BK;
            tma_load_4d(sv, &map_v, &full_bar[stage], k % p.kblock, mb * BM, k / p.kblock, 0);
          } else {
            tma_load_3d(sv, &map_v, &full_bar[stage], kb * BK, mb * BM, 0);
          }
          tma_load_2d(sv + NS * BM * BK, &map_a, &full_bar[stage], kb * BK, nb * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      int stage = 0; uint32_t phase = 0, tphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tmem_empty, tphase ^ 1);              // epilogue has drained the accumulators
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sv = smem_u32(stage_base + stage * C::kStageBytes);
          const uint64_t desc_b = make_desc(sv + NS * BM * BK);
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            const uint64_t desc_a = make_desc(sv + s * BM * BK);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              mma_i8(tmem_base + s * BN, desc_a + (uint64_t)((k * UMMA_K) >> 4),
                     desc_b + (uint64_t)((k * UMMA_K) >> 4), idesc, (kb | k) ? 1u : 0u);
          }
          mma_commit(&empty_bar[stage]);                // smem slot reusable once these MMAs finish
          if (kb == kblocks - 1) mma_commit(tmem_full); // accumulators complete
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tphase ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - kEpiWarp0;                     // == warp % 4: TMEM lane quarter
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
    uint8_t* my_stage = staging + ew * (NS * 32 * 32);
    uint32_t tphase = 0;
    double dmax = 0.0, omax = 0.0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mb, nb;
      tile_coords(p, tile, mb, nb);
      const int64_t r = (int64_t)mb * BM + ew * 32 + lane;      // row of D owned by this thread
      const int64_t n0 = (int64_t)nb * BN;
      const bool rvalid = r < p.R;
      mbar_wait(tmem_full, tphase);
      tphase ^= 1;
      tc_fence_after();

      if (MODE == SRK_I8_MID) {
        const double cin = rvalid ? row_bound(p.in_rowbound, r) / kQ : 0.0;
        const int64_t rk = r + p.diag_offset;             // column of A8 matching this row
        for (int jc = 0; jc < BN; jc += 32) {
          // per-lane output scale for column n0+jc+lane, broadcast by shuffle below
          const int64_t jl = n0 + jc + lane;
          double oscale = 0.0;
          if (jl < p.N) { const double b = row_bound(p.out_rowbound, jl); oscale = b > 0.0 ? kQ / b : 0.0; }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (jc + h * 16 >= BN) break;
            uint32_t a[NS][16];
#pragma unroll
            for (int s = 0; s < NS; ++s) tmem_ld16(lane_base + s * BN + jc + h * 16, a[s]);
            tmem_wait_ld();
            if (jc + h * 16 + 16 >= BN) {                 // last read of this tile: free the TMEM
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tmem_empty);
            }
#pragma unroll
            for (int x = 0; x < 16; ++x) {
              const int jj = h * 16 + x;
              const int64_t j = n0 + jc + jj;
              const double sc = __shfl_sync(0xffffffffu, oscale, jj);
              double u = combine<NS>(a, x) * cin;
              if (p.unit_diag && rvalid && j < p.N) u += (double)p.A8[j * p.lda + rk];
              double q = rint(u * sc);
              if (!(q > 0.0)) q = 0.0;
              if (q > kQ - 1.0) q = kQ - 1.0;
              const unsigned long long qi = rvalid ? (unsigned long long)q : 0ull;
#pragma unroll
              for (int s = 0; s < NS; ++s)
                my_stage[(s * 32 + jj) * 32 + lane] = (uint8_t)(qi >> (8 * (NS - 1 - s)));
            }
          }
          __syncwarp();
          // transposed write-out: 32 output rows (j) x 32 bytes (this warp's r range) per plane
          const int64_t col0 = (int64_t)mb * BM + ew * 32;
#pragma unroll
          for (int s = 0; s < NS; ++s) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int id = lane + 32 * h;
              const int jj = id >> 1, half = id & 1;
              const int64_t j = n0 + jc + jj;
              if (jc + jj < BN && j < p.N && col0 + half * 16 + 16 <= p.ld_outp) {
                const uint4 w = *reinterpret_cast<const uint4*>(my_stage + (s * 32 + jj) * 32 + half * 16);
                *reinterpret_cast<uint4*>(p.out_planes + s * p.out_plane_stride + j * p.ld_outp + col0 + half * 16) = w;
              }
            }
          }
          __syncwarp();
        }
      } else {
        // FINAL / COUNTS: thread owns row r, walks the BN columns 16 at a time
        double rowc = 0.0, oscale = 0.0;
        if (MODE == SRK_I8_FINAL && rvalid) {
          rowc = row_bound(p.in_rowbound, r) / kQ * p.g_row[r] * p.epi.coef;
          if (p.out_planes) { const double b = row_bound(p.out_rowbound, r); oscale = b > 0.0 ? kQ / b : 0.0; }
        }
        const int64_t rg = r + p.diag_offset;             // global column index of the diagonal
        for (int jc = 0; jc < BN; jc += 16) {
          uint32_t a[NS][16];
#pragma unroll
          for (int s = 0; s < NS; ++s) tmem_ld16(lane_base + s * BN + jc, a[s]);
          tmem_wait_ld();
          if (jc + 16 >= BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
          }
          const int64_t j0 = n0 + jc;
          if (!rvalid || j0 >= p.N) continue;
          if (MODE == SRK_I8_COUNTS) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int x = 0; x < 16; ++x) {
              const uint32_t c = (j0 + x < p.N) ? min(a[0][x], 255u) : 0u;
              w[x >> 2] |= c << (8 * (x & 3));
            }
            if (j0 + 16 <= p.ld_outp)
              *reinterpret_cast<uint4*>(p.out_planes + r * p.ld_outp + j0) = make_uint4(w[0], w[1], w[2], w[3]);
            continue;
          }
          // ---- FINAL
          uint32_t ev[4] = {0u, 0u, 0u, 0u};
          if (p.epi.evidence) {
            const uint8_t* e = p.epi.evidence + r * p.epi.ld_evidence + j0;
            if (j0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(e) & 15) == 0)) {
              const uint4 t = *reinterpret_cast<const uint4*>(e);
              ev[0] = t.x; ev[1] = t.y; ev[2] = t.z; ev[3] = t.w;
            } else {
#pragma unroll
              for (int x = 0; x < 16; ++x)
                if (j0 + x < p.N) ev[x >> 2] |= (uint32_t)e[x] << (8 * (x & 3));
            }
          }
          uint32_t w[NS][4];
#pragma unroll
          for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int x = 0; x < 4; ++x) w[s][x] = 0u;
          double v[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const int64_t j = j0 + x;
            double val = 0.0;
            if (j < p.N) {
              val = combine<NS>(a, x) * rowc * p.g_col[j];
              if (p.epi.evidence) val *= evidence_factor((ev[x >> 2] >> (8 * (x & 3))) & 0xffu);
              if (p.epi.prior) val = (1.0 - p.epi.lambda) * val + p.epi.lambda * p.epi.prior[r * p.epi.ld_prior + j];
              if (j == rg) val = 1.0; else if (val > omax) omax = val;
              if (p.epi.s_old) {
                const double d = fabs(val - p.epi.s_old[r * p.epi.ld_s_old + j]);
                if (d > dmax) dmax = d;
              }
              if (p.out_planes && j != rg) {
                double q = rint(val * oscale);
                if (!(q > 0.0)) q = 0.0;
                if (q > kQ - 1.0) q = kQ - 1.0;
                const unsigned long long qi = (unsigned long long)q;
#pragma unroll
                for (int s = 0; s < NS; ++s)
                  w[s][x >> 2] |= (uint32_t)((qi >> (8 * (NS - 1 - s))) & 0xffull) << (8 * (x & 3));
              }
            }
            v[x] = val;
          }
          double* orow = p.out_f64 + r * p.ld_out + j0;
          if (j0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0)) {
#pragma unroll
            for (int x = 0; x < 16; x += 2) *reinterpret_cast<double2*>(orow + x) = make_double2(v[x], v[x + 1]);
          } else {
#pragma unroll
            for (int x = 0; x < 16; ++x)
              if (j0 + x < p.N) orow[x] = v[x];
          }
          if (p.out_planes && j0 + 16 <= p.ld_outp) {
#pragma unroll
            for (int s = 0; s < NS; ++s)
              *reinterpret_cast<uint4*>(p.out_planes + s * p.out_plane_stride + r * p.ld_outp + j0) =
                  make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
          }
        }
      }
    }
    if (MODE == SRK_I8_FINAL) {
      dmax = warp_max(dmax);
      omax = warp_max(omax);
      if (lane == 0) {
        if (p.maxdiff && dmax > 0.0) atomic_max_nonneg(p.maxdiff, dmax);
        if (p.maxoff && omax > 0.0) atomic_max_nonneg(p.maxoff, omax);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
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

// u8 tensor map with 128B swizzle.  Dimensions (fastest first): columns, rows, [k-blocks], [planes];
// the box is BK columns x box_rows rows x 1 k-block x box_planes planes, so that the planes of one
// BM x BK operand tile land back to back in shared memory.
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

template <int NS, int BN, int MODE>
static int launch(const srk_i8_args& a, cudaStream_t st) {
  using C = Cfg<NS, BN>;
  CUtensorMap map_v, map_a;
  int rc;
  if (a.in_kblock > 0) {
    cuuint64_t dims[4] = {(cuuint64_t)a.in_kblock, (cuuint64_t)a.R, (cuuint64_t)(a.K / a.in_kblock), (cuuint64_t)NS};
    cuuint64_t strides[3] = {(cuuint64_t)a.ld_in, (cuuint64_t)a.in_kblock_stride, (cuuint64_t)a.in_plane_stride};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)BM, 1, (cuuint32_t)NS};
    rc = make_map(&map_v, a.in_planes, 4, dims, strides, box);
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)a.K, (cuuint64_t)a.R, (cuuint64_t)NS};
    cuuint64_t strides[2] = {(cuuint64_t)a.ld_in, (cuuint64_t)a.in_plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, (cuuint32_t)NS};
    rc = make_map(&map_v, a.in_planes, 3, dims, strides, box);
  }
  if (rc) return rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.N};
    cuuint64_t strides[1] = {(cuuint64_t)a.lda};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    rc = make_map(&map_a, a.A8, 2, dims, strides, box);
  }
  if (rc) return rc;
  Params p;
  memset(&p, 0, sizeof(p));
  p.mode = MODE; p.unit_diag = a.unit_diag;
  p.R = a.R; p.N = a.N; p.K = a.K;
  p.in_rowbound = a.in_rowbound;
  p.A8 = a.A8; p.lda = a.lda; p.diag_offset = a.diag_offset;
  p.g_row = a.g_row; p.g_col = a.g_col;
  p.out_f64 = a.out_f64; p.ld_out = a.ld_out;
  p.out_planes = a.out_planes; p.ld_outp = a.ld_outp; p.out_plane_stride = a.out_plane_stride;
  p.out_rowbound = a.out_rowbound;
  p.epi = to_dev(a.epi);
  p.maxdiff = a.epi.maxdiff; p.maxoff = a.epi.maxoff;
  p.tiles_m = (int)((a.R + BM - 1) / BM);
  p.tiles_n = (int)((a.N + BN - 1) / BN);
  p.group_m = 8;
  p.kblock = (int)a.in_kblock;
  int dev = 0, sms = 0;
  SRK_CUDA_OK(cudaGetDevice(&dev));
  SRK_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  const int grid = (int)(tiles < sms ? tiles : sms);
  auto kern = i8_half_kernel<NS, BN, MODE>;
  SRK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(map_v, map_a, p);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

}  // namespace i8
}  // namespace srk

using namespace srk;

extern "C" int srk_i8_supported(void) {
  int cc = srk_device_cc();
  return cc >= 100 && cc < 103 ? 1 : 0;     // kind::i8 exists on sm_100a/sm_101a only
}

extern "C" int srk_i8_half(const srk_i8_args* a, void* stream) {
  SRK_REQUIRE(a, "null args");
  SRK_REQUIRE(a->in_planes && a->A8, "null operand");
  SRK_REQUIRE(a->R > 0 && a->N > 0 && a->K > 0, "empty problem");
  SRK_REQUIRE(a->ld_in % 16 == 0 && a->lda % 16 == 0 && a->in_plane_stride % 16 == 0, "operand strides must be multiples of 16");
  SRK_REQUIRE(((uintptr_t)a->in_planes % 16) == 0 && ((uintptr_t)a->A8 % 16) == 0, "operands must be 16-byte aligned");
  SRK_REQUIRE(a->lda >= a->K, "lda smaller than K");
  if (a->in_kblock > 0) {
    SRK_REQUIRE(a->in_kblock % 128 == 0 && a->K % a->in_kblock == 0 && a->ld_in >= a->in_kblock &&
                    a->in_kblock_stride % 16 == 0, "K-blocked operand: in_kblock must be a multiple of 128 dividing K");
  } else {
    SRK_REQUIRE(a->ld_in >= a->K, "ld_in smaller than K");
  }
  SRK_REQUIRE(a->K < (1ll << 22), "K too large for exact int32 accumulation");
  if (!srk_i8_supported()) return fail(SRK_ERR_UNSUPPORTED, "%s", "tcgen05 kind::i8 needs an sm_100 device");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->mode == SRK_I8_COUNTS) {
    SRK_REQUIRE(a->ns == 1 && a->out_planes && a->ld_outp % 16 == 0 && a->ld_outp >= a->N, "COUNTS needs ns=1 and a uint8 output");
    SRK_REQUIRE(((uintptr_t)a->out_planes % 16) == 0, "output must be 16-byte aligned");
    return i8::launch<1, 256, SRK_I8_COUNTS>(*a, st);
  }
  if (a->mode == SRK_I8_MID) {
    SRK_REQUIRE(a->out_planes, "MID needs output planes");
    SRK_REQUIRE(a->ld_outp % 16 == 0 && a->out_plane_stride % 16 == 0 && ((uintptr_t)a->out_planes % 16) == 0,
                "output planes must be 16-byte aligned with ld multiple of 16");
    SRK_REQUIRE(a->ld_outp >= a->R, "MID output is transposed: ld_outp >= R");
    SRK_REQUIRE(!a->unit_diag || a->R + a->diag_offset <= a->K, "unit diagonal outside A8");
    switch (a->ns) {
      case 2: return i8::launch<2, 256, SRK_I8_MID>(*a, st);
      case 3: return i8::launch<3, 160, SRK_I8_MID>(*a, st);
      case 4: return i8::launch<4, 128, SRK_I8_MID>(*a, st);
    }
    return fail(SRK_ERR_INVALID, "invalid argument: %s", "ns must be 2, 3 or 4");
  }
  if (a->mode == SRK_I8_FINAL) {
    SRK_REQUIRE(a->out_f64 && a->g_row && a->g_col, "FINAL needs out_f64, g_row, g_col");
    SRK_REQUIRE(a->ld_out >= a->N, "ld_out");
    if (a->out_planes) {
      SRK_REQUIRE(a->ld_outp % 16 == 0 && a->ld_outp >= a->N && a->out_plane_stride % 16 == 0 &&
                      ((uintptr_t)a->out_planes % 16) == 0, "output planes must be 16-byte aligned with ld multiple of 16");
    }
    switch (a->ns) {
      case 2: return i8::launch<2, 256, SRK_I8_FINAL>(*a, st);
      case 3: return i8::launch<3, 160, SRK_I8_FINAL>(*a, st);
      case 4: return i8::launch<4, 128, SRK_I8_FINAL>(*a, st);
    }
    return fail(SRK_ERR_INVALID, "invalid argument: %s", "ns must be 2, 3 or 4");
  }
  return fail(SRK_ERR_INVALID, "invalid argument: %s", "unknown mode");
}
