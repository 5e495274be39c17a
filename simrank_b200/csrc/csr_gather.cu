// CSR half-products: the neighbour rows of X are gathered into shared memory by TMA
// (cp.async.bulk.tensor.2d.tile::gather4: FOUR rows of X per instruction, completion counted on an
// mbarrier) and summed by the warp that owns the graph row.  Two arithmetic modes share the kernel:
//
//   f64  OUT[c, i] = g[i] * sum_{m in N(i)} X[m, c]                      exact float64 (DADD)
//   u16  D[i, c]   = sum_{m in N(i)} Xq[m, c]    Xq uint16 fixed point with one scale per COLUMN of
//        X, so the sum over m is an exact integer (deg < 65536 keeps it inside 32 bits); the scales
//        are applied once per output element in the epilogue.  4x fewer gathered bytes.
//
// Why it looks like this (DESIGN.md "K3"; numbers from profiles/r2_ncu_full_csr_f64_baseline.json and
// profiles/r2_micro_tma_gather_rate.txt).  The streamed operands are O(n^2) bytes, the gather is
// nnz * n * sizeof(element): 550 GB per half-product at BASELINE cfg4 in float64, served by L2 (all
// CTAs of a grid column share one column panel of X, which is what keeps it there).  The previous
// kernel (4 __ldg row segments in flight per warp) moved 549 GB over the L2->SM crossbar in 35.4 ms
// = 15.5 TB/s with an L1 hit rate of 0: the half-product sits on the L2 bandwidth roof (51 B/clk per
// SM), not on HBM (27 GB of DRAM traffic).  So (1) the bytes are cut: uint16 planes of the same
// row-max-scaled fixed point the tensor-core path uses, and a symmetric second half; (2) the gather
// is taken off the register file and the LSU: a TMA instruction costs ~30-40 issue cycles per SM
// whatever it moves (a 1-D bulk copy of 256 B .. 1 KB tops out at 9 .. 25 B/clk per SM), so rows are
// fetched four at a time in 1 KB segments -- 4 KB per instruction, ~100 B/clk per SM, twice the L2
// roof -- into a per-warp ring; the lanes read the landed rows with conflict-free 16-byte
// ld.shared and add them up.
//
// CTA tile: TI = 32 graph rows x TC columns of X.  The rows are dealt to the 8 warps in contiguous
// runs of about equal EDGE count (a hub row gets a warp to itself), and a warp walks its rows'
// neighbour lists four entries at a time; the TMA issue runs kDepth instructions ahead of the
// consumption, across row boundaries.
//
// Epilogues
//   FIRST            T = (G X)^T: per finished row the values go to a shared-memory tile, the CTA
//                    writes the tile transposed (rows of OUT are contiguous in i).  u16: re-quantised
//                    with the bound of output column i (a per-node vector times a scalar).
//   FINAL            the same transposed store with the fused SimRank epilogue (srk_epilogue).
//   FINAL symmetric  square problems whose result is symmetric (no prior): only tiles that contain
//                    an element c >= i are computed, the epilogue runs in the row-major orientation
//                    straight from the accumulators (vector loads of S_old / counts, vector store of
//                    row i) and every off-diagonal value is ALSO stored at (c, i): each unordered pair
//                    is computed once, S stays bit-exactly symmetric, the second half gathers
//                    about half of the bytes.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace srk {
namespace gat {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTI = 32;                     // graph rows per CTA (the transposed second half takes 16, see Smem)
constexpr int kTIAccum = 8;                 // MODE_ACCUM: pieces of neighbour lists per CTA (equal sizes: one per warp)
constexpr int kRowsPerOp = 4;               // rows of X per TMA instruction (tile::gather4)

constexpr int MODE_FIRST = 0, MODE_FINAL = 1, MODE_FINAL_SYM = 2, MODE_ACCUM = 3;

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "GAT_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra GAT_DONE;\n\t"
      "bra GAT_WAIT;\n\t"
      "GAT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// rows r0..r3 of the 2-D tensor, box-width columns from column `col`, into 4 consecutive row
// segments at dst; 4 * segment bytes are counted on `bar`
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, uint64_t* bar, uint64_t pol, int col,
                                            int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct Params {
  // neighbour list of row r: indices[rowbeg[r] .. rowend[r]).  Normally indptr / indptr + 1; with hub rows
  // split off (srk_csr_args.row_ptr_*), the hubs' lists are empty here and their sums arrive through
  // `accum`; in MODE_ACCUM the "rows" are the chunks of the hub rows.
  const int64_t* rowbeg; const int64_t* rowend; const int32_t* indices; const double* g;
  uint32_t* accum; int64_t ld_accum; const int32_t* accum_slot;   // u16: partial sums of the hub rows, [slots][ld] u32
  int64_t row_begin, row_end;
  const void* X; int64_t ldx, L;
  void* OUT; int64_t ldo;
  srk_rowbound in_unit;        // u16: value of one unit of column c of X
  srk_rowbound out_bound;      // u16 FIRST: bound of output column i (one unit = bound / qmax)
  double qmax;                 // u16: largest fixed-point value (65535, less when a row degree exceeds 65536)
  const double* g_col;         // u16 FINAL: row factor of output row r (= column of X)
  const void* counts; int64_t ld_counts; int counts32, add_counts, use_evidence;
  EpilogueDev epi; double* maxdiff; double* maxoff;
  int tma;                     // rows of X are 16-byte aligned: TMA gather; else plain loads
  int vec_aligned;             // OUT, S_old and counts are 16-byte aligned (vector epilogue of the symmetric half)
  int upper_only;              // ACCUM for a symmetric second half: a piece of row i skips the panels left of column i
  int flags;                   // SRK_CSR_FLAGS (A/B profiling): 1 = default L2 policy for the gather
  int64_t tiles_x;             // row tiles of the problem (symmetric second half: triangular 1-D grid)
};

// uint32 -> double and double -> nearest integer without the conversion instructions (I2F.F64 / FRND / F2I issue
// at a quarter of the rate of DADD, profiles/r1_micro_fp64_rate.txt): 2^52 + w has w in its low mantissa word.
__device__ __forceinline__ double u32_to_double(uint32_t w) {
  return __hiloint2double(0x43300000, (int)w) - 4503599627370496.0;
}
// rint(x) clipped to [0, qmax] as uint32, for finite or NaN x (NaN -> 0); qmax is an integer below 2^32, so
// clipping first gives what rint-then-clip gives
__device__ __forceinline__ uint32_t rint_clip_u32(double x, double qmax) {
  x = fmin(fmax(x, 0.0), qmax);
  return (uint32_t)__double2loint(x + 4503599627370496.0);
}

__device__ __forceinline__ uint32_t load_count(const void* base, int64_t idx, int c32) {
  return c32 ? reinterpret_cast<const uint32_t*>(base)[idx] : (uint32_t)reinterpret_cast<const uint16_t*>(base)[idx];
}

// Accumulators of one graph row.  A row segment is kSeg bytes = G groups of 512 B; lane l reads the
// 16 bytes at 16 l of every group (one conflict-free ld.shared.v4 per group), so it owns
// kVec = 16 / sizeof(element) CONSECUTIVE columns per group: column(j) = (j / kVec) * (512 / size) +
// kVec * l + j % kVec.
template <typename E, int TC>
struct Acc;

template <int TC>
struct Acc<double, TC> {
  static constexpr int kGroups = TC * 8 / 512, kVec = 2, kCols = kGroups * kVec, kGroupCols = 64;
  double v[kCols];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < kCols; ++k) v[k] = 0.0;
  }
  __device__ __forceinline__ void add_smem(const uint8_t* seg, int lane) {
#pragma unroll
    for (int gq = 0; gq < kGroups; ++gq) {
      const double2 x = *reinterpret_cast<const double2*>(seg + gq * 512 + lane * 16);
      v[2 * gq] += x.x;
      v[2 * gq + 1] += x.y;
    }
  }
  __device__ __forceinline__ void add_global(const void* row, int lane, int64_t valid) {
    const double* s = reinterpret_cast<const double*>(row);
#pragma unroll
    for (int j = 0; j < kCols; ++j) {
      const int c = (j / kVec) * kGroupCols + kVec * lane + j % kVec;
      v[j] += c < valid ? __ldg(s + c) : 0.0;
    }
  }
  __device__ __forceinline__ double val(int j) const { return v[j]; }
  __device__ __forceinline__ uint32_t raw(int) const { return 0u; }
  __device__ __forceinline__ void load_sums(const uint32_t*, int) {}
};

// uint16: the low halves of the 32-bit words are not masked out per element: aw accumulates the
// whole words modulo 2^32 and ah the high halves, so sum(lo) = aw - (ah << 16) (mod 2^32), which is
// exact because sum(lo) < 2^32.
template <int TC>
struct Acc<uint16_t, TC> {
  static constexpr int kGroups = TC * 2 / 512, kVec = 8, kCols = kGroups * kVec, kGroupCols = 256;
  uint32_t aw[kCols / 2], ah[kCols / 2];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int w = 0; w < kCols / 2; ++w) aw[w] = ah[w] = 0u;
  }
  __device__ __forceinline__ void add_word(int w, uint32_t x) {
    aw[w] += x;
    ah[w] += x >> 16;
  }
  __device__ __forceinline__ void add_smem(const uint8_t* seg, int lane) {
#pragma unroll
    for (int gq = 0; gq < kGroups; ++gq) {
      const uint4 x = *reinterpret_cast<const uint4*>(seg + gq * 512 + lane * 16);
      add_word(4 * gq, x.x);
      add_word(4 * gq + 1, x.y);
      add_word(4 * gq + 2, x.z);
      add_word(4 * gq + 3, x.w);
    }
  }
  __device__ __forceinline__ void add_global(const void* row, int lane, int64_t valid) {
    const uint16_t* s = reinterpret_cast<const uint16_t*>(row);
#pragma unroll
    for (int w = 0; w < kCols / 2; ++w) {
      const int c = (w / 4) * kGroupCols + kVec * lane + 2 * (w % 4);
      const uint32_t lo = c < valid ? (uint32_t)__ldg(s + c) : 0u, hi = c + 1 < valid ? (uint32_t)__ldg(s + c + 1) : 0u;
      aw[w] += lo | (hi << 16);
      ah[w] += hi;
    }
  }
  __device__ __forceinline__ uint32_t raw(int j) const { return (j & 1) ? ah[j >> 1] : aw[j >> 1] - (ah[j >> 1] << 16); }
  __device__ __forceinline__ double val(int j) const { return u32_to_double(raw(j)); }
  // start from per-column sums computed elsewhere (hub rows): s[c] for the panel's columns, 32-byte aligned
  __device__ __forceinline__ void load_sums(const uint32_t* s, int lane) {
#pragma unroll
    for (int gq = 0; gq < kGroups; ++gq) {
      const uint4* q = reinterpret_cast<const uint4*>(s + gq * kGroupCols + kVec * lane);
      const uint4 lo = q[0], hi = q[1];
      const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) { ah[4 * gq + w] = v[2 * w + 1]; aw[4 * gq + w] = v[2 * w] + (v[2 * w + 1] << 16); }
    }
  }
};

// One element of the fixed-point second half, x = g_i g_r (D unit_r + counts) coef evidence.  Every rounding
// is spelled out (no FMA contraction): the transposed second half and the FINISH pass (csr_finish_kernel)
// must agree bit for bit, whatever the compiler does around them.
__device__ __forceinline__ double final_value_u16(double gi, double gc, double sum, double fu, double cnt_term,
                                                  double coef, double evf) {
  const double v = __dmul_rn(__dmul_rn(gi, gc), __dadd_rn(__dmul_rn(sum, fu), cnt_term));
  return __dmul_rn(__dmul_rn(v, coef), evf);
}

template <typename E, int TC>
__device__ __forceinline__ int col_of(int j, int lane) {
  return (j / Acc<E, TC>::kVec) * Acc<E, TC>::kGroupCols + Acc<E, TC>::kVec * lane + j % Acc<E, TC>::kVec;
}

template <typename E, int MODE>
struct TileElem { typedef double type; };
template <>
struct TileElem<uint16_t, MODE_FIRST> { typedef uint16_t type; };
template <>
struct TileElem<uint16_t, MODE_FINAL> { typedef uint32_t type; };

template <typename E, int TC, int MODE>
struct Smem {
  typedef typename TileElem<E, MODE>::type TileT;
  // The transposed second half keeps 4 or 8 bytes per element of its tile in shared memory: with 16
  // rows per CTA the tile of a full-width panel (1 KB segments) still leaves room for two CTAs per SM,
  // and a row of the result is still a whole 128-byte line.
  static constexpr int TI = MODE == MODE_FINAL ? 16 : (MODE == MODE_ACCUM ? kTIAccum : kTI);
  static constexpr int kSeg = TC * (int)sizeof(E);                   // bytes of one row segment (<= 1024)
  static constexpr int kSlot = kRowsPerOp * kSeg;
  // TMA instructions in flight per warp: as many as shared memory allows with two CTAs per SM
  static constexpr int kDepth = kSlot >= 4096 ? ((MODE == MODE_FINAL_SYM || MODE == MODE_ACCUM) ? 3 : 2) : 4;
  static constexpr int kRing = kWarps * kDepth * kSlot;
  static constexpr int kBars = kWarps * kDepth * 8;
  // FINAL symmetric: per-column factors of the panel, fa = coef g_col unit, fb = coef g_col
  static constexpr int kFac = (MODE == MODE_FINAL_SYM && sizeof(E) == 2) ? TC * 16 : 0;
  // transposed-store tile [TC][kPitch]: odd pitch in 32-bit words where the element size allows
  static constexpr int kPitch = sizeof(TileT) == 2 ? TI + 2 : TI + 1;
  static constexpr int kRowsPerPass = 32 / TI;                        // tile rows a warp stores per pass of the transposed store
  static constexpr int kTile = (MODE == MODE_FINAL_SYM || MODE == MODE_ACCUM) ? 0 : TC * kPitch * (int)sizeof(TileT);
  static constexpr int kBytes = kRing + kBars + kFac + kTile + 128;
  static_assert(kSeg % 512 == 0 && kSeg <= 1024, "row segments are 512 B or 1 KB (TMA box <= 256 elements of 4 B)");
};

template <typename E, int TC, int MODE>
__global__ void __launch_bounds__(kThreads, 2)   // shared memory allows two CTAs per SM in every mode
csr_gather_kernel(const __grid_constant__ CUtensorMap map_x, const Params p) {
  typedef Smem<E, TC, MODE> SM;
  typedef typename SM::TileT TileT;
  typedef Acc<E, TC> A;
  constexpr bool kU16 = sizeof(E) == 2;
  constexpr int kDepth = SM::kDepth;
  constexpr int TI = SM::TI;
  extern __shared__ uint8_t smem_raw[];
  __shared__ double red[2][kWarps];
  __shared__ int next_row;                    // rows of the tile are claimed by the warps as they go
  __shared__ uint8_t fifo[kWarps][kTI];       // rows a warp has claimed, in order (issue side -> consume side)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kRing);
  double2* fac = reinterpret_cast<double2*>(smem + SM::kRing + SM::kBars);
  TileT* tile = reinterpret_cast<TileT*>(smem + SM::kRing + SM::kBars + SM::kFac);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t bx = blockIdx.x, by = blockIdx.y;
  if (MODE == MODE_FINAL_SYM) {
    // 1-D grid over the tiles that contain an element c >= i: panel y needs the row tiles
    // x < min(tiles_x, (y + 1) * TC / TI); tiles are numbered panel by panel
    constexpr int64_t kPer = TC / TI;                           // row tiles added per panel
    const int64_t t = blockIdx.x, full = p.tiles_x / kPer;      // panels before the count saturates at tiles_x
    const int64_t tri = kPer * full * (full + 1) / 2;           // tiles in panels 0 .. full - 1
    if (t < tri) {
      by = (int64_t)((sqrt(1.0 + 8.0 * (double)t / (double)kPer) - 1.0) * 0.5);
      while (kPer * by * (by + 1) / 2 > t) --by;
      while (kPer * (by + 1) * (by + 2) / 2 <= t) ++by;
      bx = t - kPer * by * (by + 1) / 2;
    } else {
      by = full + (t - tri) / p.tiles_x;
      bx = (t - tri) % p.tiles_x;
    }
  }
  const int64_t i0 = p.row_begin + bx * TI;
  const int64_t c0 = by * TC;
  const int rows_here = (int)min((int64_t)TI, p.row_end - i0);

  uint8_t* wring = smem + (size_t)warp * kDepth * SM::kSlot;
  uint64_t* wbar = bars + warp * kDepth;
  if (p.tma && lane < kDepth) mbar_init(&wbar[lane], 1);
  if (p.tma) fence_barrier_init();
  if (threadIdx.x == 0) next_row = 0;
  if (MODE == MODE_FINAL && i0 < p.row_end) {
    // the transposed epilogue reads TI consecutive elements of S_old (and counts) for each of the TC
    // output rows of this tile once the gather is done: pull those lines into L2 now
    for (int c = threadIdx.x; c < TC; c += kThreads) {
      const int64_t r = c0 + c;
      if (r >= p.L) break;
      if (p.epi.s_old) prefetch_l2(p.epi.s_old + r * p.epi.ld_s_old + i0);
      if (p.counts) prefetch_l2(reinterpret_cast<const uint8_t*>(p.counts) + (r * p.ld_counts + i0) * (p.counts32 ? 4 : 2));
    }
  }
  if (MODE == MODE_FINAL_SYM && kU16) {
    for (int c = threadIdx.x; c < TC; c += kThreads) {
      const int64_t cc = c0 + c;
      const double gc = cc < p.L ? p.g_col[cc] * p.epi.coef : 0.0;
      fac[c] = make_double2(cc < p.L ? gc * row_bound(p.in_unit, cc) : 0.0, gc);
    }
  }
  __syncthreads();

  double dmax = 0.0, omax = 0.0;
  {
    const int64_t valid = min((int64_t)TC, p.ldx - c0);               // columns of X this panel can read
    const uint8_t* xbase = reinterpret_cast<const uint8_t*>(p.X) + c0 * (int64_t)sizeof(E);
    const int64_t pitch = p.ldx * (int64_t)sizeof(E);
    const uint64_t pol = (p.flags & 1) ? policy_evict_normal() : policy_evict_last();
    const int col32 = (int)(c0 * (int64_t)sizeof(E) / 4);             // TMA coordinate in 4-byte elements

    // ---- issue side.  Rows are claimed from the tile's counter when the previous one runs out (a warp
    // stuck on a hub row claims nothing while the others share the rest) and pushed to the warp's FIFO
    // for the consume side; the neighbour indices of the NEXT instruction are loaded one step early,
    // so the refill of a slot does not wait for them.
    int wi = 0, ri = 0;                                               // FIFO write / read positions
    int64_t ip = 0, iend = 0;
    int pre_my = 0;
    bool have_pre = false;
    unsigned issued = 0;
    auto advance = [&]() {
      have_pre = false;
      while (ip >= iend) {
        int r = 0;
        if (lane == 0) r = atomicAdd(&next_row, 1);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= rows_here) return;
        if (MODE == MODE_ACCUM && p.upper_only) {                       // nothing of this panel lies at or right of
          const int sl = p.accum_slot[i0 + r];                          // the diagonal of the piece's row (slot = row)
          if ((int64_t)(sl >= 0 ? sl : -sl - 1) >= c0 + TC) continue;
        }
        if (lane == 0) fifo[warp][wi] = (uint8_t)r;
        ++wi;
        ip = p.rowbeg[i0 + r];
        iend = p.rowend[i0 + r];
      }
      const int cnt = (int)min((int64_t)kRowsPerOp, iend - ip);
      pre_my = p.indices[ip + min(lane & 3, cnt - 1)];                // a short group repeats its last row
      ip += kRowsPerOp;
      have_pre = true;
    };
    auto issue_next = [&]() {
      if (!have_pre) return;
      const int m0 = __shfl_sync(0xffffffffu, pre_my, 0), m1 = __shfl_sync(0xffffffffu, pre_my, 1);
      const int m2 = __shfl_sync(0xffffffffu, pre_my, 2), m3 = __shfl_sync(0xffffffffu, pre_my, 3);
      if (lane == 0) {
        const int slot = (int)(issued % kDepth);
        mbar_expect_tx(&wbar[slot], SM::kSlot);
        tma_gather4(wring + slot * SM::kSlot, &map_x, &wbar[slot], pol, col32, m0, m1, m2, m3);
      }
      ++issued;
      advance();
    };
    advance();
    if (p.tma) {
#pragma unroll 1
      for (int d = 0; d < kDepth; ++d) issue_next();
    }

    A acc;
    unsigned consumed = 0;
    while (true) {
      __syncwarp();                                                   // FIFO entries written by lane 0
      if (ri == wi) break;                                            // every claimed row is done, none left to claim
      const int64_t row = i0 + fifo[warp][ri];
      ++ri;
      const int64_t rbeg = p.rowbeg[row], rend = p.rowend[row];
      if (MODE == MODE_FINAL_SYM && p.epi.s_old) {
        // the epilogue of this row reads S_old (and the counts) once the gather is done: pull its
        // lines into L2 now, so that it sees L2 latency instead of DRAM latency
#pragma unroll
        for (int gq = 0; gq < A::kGroups; ++gq) {
          const int64_t c = c0 + gq * A::kGroupCols + A::kVec * lane;
          if (c < p.L && c + A::kVec > row) {
            prefetch_l2(p.epi.s_old + row * p.epi.ld_s_old + c);
            if (A::kVec > 4 && c + 4 < p.L) prefetch_l2(p.epi.s_old + row * p.epi.ld_s_old + c + 4);
            if (p.counts && (lane & 1) == 0)                      // 32 B of counts cover two lanes' columns
              prefetch_l2(reinterpret_cast<const uint8_t*>(p.counts) + (row * p.ld_counts + c) * (p.counts32 ? 4 : 2));
          }
        }
      }
      acc.clear();
      if (MODE != MODE_ACCUM && kU16 && p.accum_slot) {
        const int slot = p.accum_slot[row];                       // a hub row: its chunks were summed beforehand
        if (slot >= 0) acc.load_sums(p.accum + (int64_t)slot * p.ld_accum + c0, lane);
      }
      if (p.tma) {
        for (int64_t e = rbeg; e < rend; e += kRowsPerOp) {
          const int cnt = (int)min((int64_t)kRowsPerOp, rend - e);
          const int slot = (int)(consumed % kDepth);
          mbar_wait(&wbar[slot], (consumed / kDepth) & 1u);
          const uint8_t* sp = wring + slot * SM::kSlot;
          acc.add_smem(sp, lane);
          if (cnt > 1) acc.add_smem(sp + SM::kSeg, lane);
          if (cnt > 2) acc.add_smem(sp + 2 * SM::kSeg, lane);
          if (cnt > 3) acc.add_smem(sp + 3 * SM::kSeg, lane);
          ++consumed;
          __syncwarp();                                             // every lane has read the slot
          issue_next();                                             // refills it (kDepth steps ahead)
        }
      } else {
        for (int64_t e = rbeg; e < rend; e += 32) {
          const int cnt = (int)min((int64_t)32, rend - e);
          const int my = lane < cnt ? p.indices[e + lane] : 0;
          for (int t = 0; t < cnt; ++t)
            acc.add_global(xbase + (int64_t)__shfl_sync(0xffffffffu, my, t) * pitch, lane, valid);
        }
        ip = iend;                                                  // plain loads: the issue side only claims rows
        advance();
      }

      // ------------------------------------------------------------------ row `row` is complete
      const int il = (int)(row - i0);
      if (MODE == MODE_ACCUM) {
        // a piece of a neighbour list: add its column sums to the row's slot (integer adds: the order does
        // not matter); a piece that IS the whole list (slot stored as -slot - 1) simply stores them
        const int sl = p.accum_slot[row];
        uint32_t* dst = p.accum + (int64_t)(sl >= 0 ? sl : -sl - 1) * p.ld_accum + c0;
        if (sl >= 0) {
#pragma unroll
          for (int j = 0; j < A::kCols; ++j) {
            const int cl = col_of<E, TC>(j, lane);
            if (c0 + cl < p.L) atomicAdd(dst + cl, acc.raw(j));
          }
        } else {
#pragma unroll
          for (int gq = 0; gq < A::kGroups; ++gq) {                   // 32 bytes per lane and group; ld_accum covers the panel
            const int j0 = gq * A::kVec;
            uint4* q = reinterpret_cast<uint4*>(dst + gq * A::kGroupCols + A::kVec * lane);
            q[0] = make_uint4(acc.raw(j0), acc.raw(j0 + 1), acc.raw(j0 + 2), acc.raw(j0 + 3));
            if (A::kVec > 4) q[1] = make_uint4(acc.raw(j0 + 4), acc.raw(j0 + 5), acc.raw(j0 + 6), acc.raw(j0 + 7));
          }
        }
      } else if (MODE == MODE_FIRST) {
        if (kU16) {
          const double bo = row_bound(p.out_bound, row);
          const double inv = bo > 0.0 ? p.qmax / bo : 0.0;
#pragma unroll
          for (int j = 0; j < A::kCols; ++j) {
            const int cl = col_of<E, TC>(j, lane);
            const int64_t c = c0 + cl;
            const uint32_t q = c < p.L ? rint_clip_u32(acc.val(j) * row_bound(p.in_unit, c) * inv, p.qmax) : 0u;
            tile[cl * SM::kPitch + il] = (TileT)q;
          }
        } else {
          const double gi = p.g[row];
#pragma unroll
          for (int j = 0; j < A::kCols; ++j) tile[col_of<E, TC>(j, lane) * SM::kPitch + il] = (TileT)(acc.val(j) * gi);
        }
      } else if (MODE == MODE_FINAL) {
        if (kU16) {
#pragma unroll
          for (int j = 0; j < A::kCols; ++j) tile[col_of<E, TC>(j, lane) * SM::kPitch + il] = (TileT)acc.val(j);
        } else {
          const double gi = p.g[row];
#pragma unroll
          for (int j = 0; j < A::kCols; ++j) tile[col_of<E, TC>(j, lane) * SM::kPitch + il] = (TileT)(acc.val(j) * gi);
        }
      } else {
        // symmetric FINAL: element (row, c) for c >= row, also stored at (c, row).  Per group the
        // lane owns kVec consecutive columns: vector loads of counts / S_old, vector store of the row.
        const double gi = kU16 ? p.g[row] : p.g[row] * p.epi.coef;
        double* orow = reinterpret_cast<double*>(p.OUT) + row * p.ldo;
#pragma unroll
        for (int gq = 0; gq < A::kGroups; ++gq) {
          const int cl0 = gq * A::kGroupCols + A::kVec * lane;
          const int64_t cb = c0 + cl0;                              // first of the kVec columns
          if (cb >= p.L || cb + A::kVec <= row) continue;           // nothing at or right of the diagonal
          const bool whole = cb >= row && cb + A::kVec <= p.L;      // every element is stored
          uint32_t cnt[A::kVec];
          double so[A::kVec], v[A::kVec];
#pragma unroll
          for (int x = 0; x < A::kVec; ++x) { cnt[x] = 0u; so[x] = 0.0; }
          const bool vec_ok = whole && p.vec_aligned && ((p.ld_counts | p.epi.ld_s_old | p.ldo) & (A::kVec - 1)) == 0;
          if (p.counts) {
            if (vec_ok && !p.counts32 && A::kVec == 8) {
              const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.counts) + row * p.ld_counts + cb);
              const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int x = 0; x < A::kVec; ++x) cnt[x] = (w[x >> 1] >> (16 * (x & 1))) & 0xffffu;
            } else {
#pragma unroll
              for (int x = 0; x < A::kVec; ++x)
                if (cb + x < p.L && cb + x >= row) cnt[x] = load_count(p.counts, row * p.ld_counts + cb + x, p.counts32);
            }
          }
          if (p.epi.s_old) {
            if (vec_ok) {
#pragma unroll
              for (int x = 0; x < A::kVec; x += 2) {
                const double2 d = *reinterpret_cast<const double2*>(p.epi.s_old + row * p.epi.ld_s_old + cb + x);
                so[x] = d.x; so[x + 1] = d.y;
              }
            } else {
#pragma unroll
              for (int x = 0; x < A::kVec; ++x)
                if (cb + x < p.L && cb + x >= row) so[x] = p.epi.s_old[row * p.epi.ld_s_old + cb + x];
            }
          }
#pragma unroll
          for (int x = 0; x < A::kVec; ++x) {
            const int64_t c = cb + x;
            const int j = gq * A::kVec + x;
            double val;
            if (kU16) {
              const double2 f = fac[cl0 + x];
              val = gi * (acc.val(j) * f.x + (p.add_counts ? (double)cnt[x] * f.y : 0.0));
            } else {
              val = acc.val(j) * gi;
            }
            if (p.use_evidence) val *= evidence_factor(cnt[x]);
            else if (p.epi.evidence && c < p.L && c >= row) val *= evidence_factor(p.epi.evidence[row * p.epi.ld_evidence + c]);
            const bool live = c < p.L && c >= row;
            if (c == row) val = 1.0; else if (live && val > omax) omax = val;
            if (live && p.epi.s_old) {
              const double d = fabs(val - so[x]);
              if (d > dmax) dmax = d;                 // NaN compares false: ignored like SimRank.py:74
            }
            v[x] = val;
          }
          if (vec_ok) {
#pragma unroll
            for (int x = 0; x < A::kVec; x += 2) __stcs(reinterpret_cast<double2*>(orow + cb + x), make_double2(v[x], v[x + 1]));
          } else {
#pragma unroll
            for (int x = 0; x < A::kVec; ++x)
              if (cb + x < p.L && cb + x >= row) __stcs(orow + cb + x, v[x]);
          }
#pragma unroll
          for (int x = 0; x < A::kVec; ++x)
            if (cb + x < p.L && cb + x > row) __stcs(reinterpret_cast<double*>(p.OUT) + (cb + x) * p.ldo + row, v[x]);
        }
      }
    }
  }

  if (MODE != MODE_FINAL_SYM && MODE != MODE_ACCUM) {
    __syncthreads();
    // transposed store: a warp pass covers kRowsPerPass tile rows (columns c of X), TI lanes each
    const int il = lane % TI, sub = lane / TI;
    const int64_t i = i0 + il;
    constexpr int kStep = kWarps * SM::kRowsPerPass;
    if (MODE == MODE_FIRST) {
      for (int cl = warp * SM::kRowsPerPass + sub; cl < TC; cl += kStep) {
        const int64_t r = c0 + cl;
        if (r >= p.L || i >= p.row_end) continue;
        if (kU16) reinterpret_cast<uint16_t*>(p.OUT)[r * p.ldo + i] = (uint16_t)tile[cl * SM::kPitch + il];
        else __stcs(reinterpret_cast<double*>(p.OUT) + r * p.ldo + i, (double)tile[cl * SM::kPitch + il]);
      }
    } else {
      // Four passes at a time with all their loads issued before the first use (eight was tried: 80 -> 109 ms on
      // the cfg5 S1 shape, profiles/r2_csr_shapes_batch8.jsonl): one pass is a chain of
      // dependent global loads (counts, S_old -- prefetched into L2 when the CTA started), and a CTA
      // has TC / kStep = 32 .. 64 of them.
      constexpr int kBatch = 4;
      const double gi = i < p.row_end ? p.g[i] : 0.0;
      for (int cb = warp * SM::kRowsPerPass + sub; cb < TC; cb += kStep * kBatch) {
        uint32_t cnt[kBatch];
        double so[kBatch], ev[kBatch], pr[kBatch], fu[kBatch], gc[kBatch];
        bool live[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int cl = cb + b * kStep;
          const int64_t r = c0 + cl;
          live[b] = cl < TC && r < p.L && i < p.row_end;
          cnt[b] = 0u; so[b] = 0.0; ev[b] = 1.0; pr[b] = 0.0; fu[b] = 0.0; gc[b] = 0.0;
          if (!live[b]) continue;
          if (p.counts) cnt[b] = load_count(p.counts, r * p.ld_counts + i, p.counts32);
          if (p.epi.s_old) so[b] = __ldcs(p.epi.s_old + r * p.epi.ld_s_old + i);
          if (!p.use_evidence && p.epi.evidence) ev[b] = evidence_factor(__ldcs(p.epi.evidence + r * p.epi.ld_evidence + i));
          if (p.epi.prior) pr[b] = __ldcs(p.epi.prior + r * p.epi.ld_prior + i);
          if (kU16) { fu[b] = row_bound(p.in_unit, r); gc[b] = p.g_col[r]; }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          if (!live[b]) continue;
          const int cl = cb + b * kStep;
          const int64_t r = c0 + cl;
          double v;
          const double evf = p.use_evidence ? evidence_factor(cnt[b]) : ev[b];
          if (kU16)
            v = final_value_u16(gi, gc[b], (double)tile[cl * SM::kPitch + il], fu[b],
                                p.add_counts ? (double)cnt[b] : 0.0, p.epi.coef, evf);
          else
            v = (double)tile[cl * SM::kPitch + il] * p.epi.coef * evf;
          if (p.epi.prior) v = (1.0 - p.epi.lambda) * v + p.epi.lambda * pr[b];
          if (r + p.epi.diag_offset == i) v = 1.0; else if (v > omax) omax = v;
          if (p.epi.s_old) {
            const double d = fabs(v - so[b]);
            if (d > dmax) dmax = d;                // NaN compares false: ignored like SimRank.py:74
          }
          // the epilogue streams are touched once: evict-first stores keep them from pushing the
          // gathered panel of X out of L2
          __stcs(reinterpret_cast<double*>(p.OUT) + r * p.ldo + i, v);
        }
      }
    }
  }
  if (MODE != MODE_FIRST && MODE != MODE_ACCUM) {
    dmax = warp_max(dmax);
    omax = warp_max(omax);
    if (lane == 0) { red[0][warp] = dmax; red[1][warp] = omax; }
    __syncthreads();
    if (warp == 0) {
      dmax = (lane < kWarps) ? red[0][lane] : 0.0;
      omax = (lane < kWarps) ? red[1][lane] : 0.0;
      dmax = warp_max(dmax);
      omax = warp_max(omax);
      if (lane == 0) {
        if (p.maxdiff && dmax > 0.0) atomic_max_nonneg(p.maxdiff, dmax);
        if (p.maxoff && omax > 0.0) atomic_max_nonneg(p.maxoff, omax);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// SRK_CSR_FINISH: the epilogue of the second half as a streaming pass of its own.  accum[i, r] holds the
// integer sums D[i, r] of every graph row i (SRK_CSR_ACCUM over pieces that cover every list); this kernel
// lays a tile of them down transposed in shared memory and runs the srk_epilogue chain in the orientation of
// the result, OUT[r, i].  Tile = 64 graph rows x 128 result rows: both sides move 512-byte runs (accum rows
// on the way in, S_old / OUT rows on the way out), a warp keeps 8 result rows of S_old and counts in flight,
// and the small tile (33 KB) leaves room for three CTAs per SM -- this pass is pure HBM traffic (4 + 8 + 2 + 8
// bytes per element), it needs bytes in flight, nothing else.
// csr_finish_first_kernel runs as many short CTAs; the tile a CTA reads was requested from DRAM by the CTA that
// ran kPrefetchAhead tiles earlier (about one wave: CTAs are dispatched in order), so its own loads see L2
// latency: 2.81 -> 2.49 ms at cfg4.  (The same trick made the two FINISH kernels, which stream three inputs,
// 8-14 % slower -- profiles/r2_csr_shapes_prefetch.jsonl -- and is not used there.)
constexpr int kPrefetchAhead = 148 * 3;
__device__ __forceinline__ void prefetch_block(const void* base, int64_t pitch_bytes, int rows, int row_bytes) {
  const int per_row = (row_bytes + 127) / 128;
  for (int t = threadIdx.x; t < rows * per_row; t += blockDim.x)
    prefetch_l2(reinterpret_cast<const uint8_t*>(base) + (int64_t)(t / per_row) * pitch_bytes + (t % per_row) * 128);
}
constexpr int kFI = 64, kFR = 128, kFPitch = kFI + 1, kFBatch = 8;
// shared memory of csr_finish_kernel: the transposed sums, then (fast path) the S_old and counts tiles, which
// arrive by cp.async -- 80 KB per CTA requested in the first microsecond, no register held while they fly
constexpr int kFTileBytes = kFR * kFPitch * 4, kFSoBytes = kFR * kFI * 8, kFCnBytes = kFR * kFI * 2;
constexpr int kFSmemBytes = kFTileBytes + kFSoBytes + kFCnBytes;
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int kCtasPerSM>
__global__ void __launch_bounds__(kThreads, kCtasPerSM)
csr_finish_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t fin_smem[];
  uint32_t* tile = reinterpret_cast<uint32_t*>(fin_smem);
  double* so_s = reinterpret_cast<double*>(fin_smem + kFTileBytes);
  uint16_t* cn_s = reinterpret_cast<uint16_t*>(fin_smem + kFTileBytes + kFSoBytes);
  __shared__ double red[2][kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i0 = p.row_begin + (int64_t)blockIdx.x * kFI, c0 = (int64_t)blockIdx.y * kFR;
  const int rows_here = (int)min((int64_t)kFI, p.row_end - i0);
  // Fast path, CTA-uniform: a full tile of the usual operands (16-bit counts, S_old present, everything aligned for
  // 16-byte accesses, no evidence array / prior).
  const bool fast = rows_here == kFI && c0 + kFR <= p.L && p.vec_aligned && (p.ldo & 1) == 0 && (p.epi.ld_s_old & 1) == 0 &&
                    (p.ld_counts & 7) == 0 && (i0 & 7) == 0 && p.epi.s_old && p.counts && !p.counts32 &&
                    !p.epi.evidence && !p.epi.prior && !(p.flags & 16);
  if (fast) {
    // S_old[c0 .. c0+128, i0 .. i0+64) and counts, row by row: 32 / 8 chunks of 16 bytes per row
    const double* sb = p.epi.s_old + c0 * p.epi.ld_s_old + i0;
    const uint16_t* cb = reinterpret_cast<const uint16_t*>(p.counts) + c0 * p.ld_counts + i0;
#pragma unroll
    for (int k = 0; k < kFSoBytes / 16 / kThreads; ++k) {
      const int ch = threadIdx.x + kThreads * k, row = ch >> 5, c16 = ch & 31;
      cp_async16(so_s + row * kFI + 2 * c16, sb + row * p.epi.ld_s_old + 2 * c16);
    }
#pragma unroll
    for (int k = 0; k < kFCnBytes / 16 / kThreads; ++k) {
      const int ch = threadIdx.x + kThreads * k, row = ch >> 3, c16 = ch & 7;
      cp_async16(cn_s + row * kFI + 8 * c16, cb + row * p.ld_counts + 8 * c16);
    }
  }

  // ---- in: 8 graph rows per warp, 16 bytes per lane and row, all eight loads issued before the first use
  {
    uint4 v[kFI / kWarps];
#pragma unroll
    for (int q = 0; q < kFI / kWarps; ++q) {
      const int il = warp + kWarps * q;
      v[q] = make_uint4(0u, 0u, 0u, 0u);
      if (il < rows_here) v[q] = __ldcs(reinterpret_cast<const uint4*>(p.accum + (i0 + il) * p.ld_accum + c0 + 4 * lane));
    }
#pragma unroll
    for (int q = 0; q < kFI / kWarps; ++q) {
      const int il = warp + kWarps * q;
      tile[(4 * lane) * kFPitch + il] = v[q].x; tile[(4 * lane + 1) * kFPitch + il] = v[q].y;
      tile[(4 * lane + 2) * kFPitch + il] = v[q].z; tile[(4 * lane + 3) * kFPitch + il] = v[q].w;
    }
  }
  if (fast) cp_async_wait_all();
  __syncthreads();

  // ---- out: 16 result rows per warp, lane l owns the graph rows i0 + 2 l, i0 + 2 l + 1
  double dmax = 0.0, omax = 0.0;
  const int64_t i = i0 + 2 * lane;
  const bool in0 = i < p.row_end, in1 = i + 1 < p.row_end;
  const double g0 = in0 ? p.g[i] : 0.0, g1 = in1 ? p.g[i + 1] : 0.0;
  const bool vec = in1 && p.vec_aligned && ((p.ldo | p.epi.ld_s_old) & 1) == 0 && (i & 1) == 0;
  const bool cvec = in1 && !p.counts32 && (p.ld_counts & 1) == 0 && (i & 1) == 0 && (reinterpret_cast<uintptr_t>(p.counts) & 3) == 0;
  double* out = reinterpret_cast<double*>(p.OUT);
  // (The general loop below spends more instructions on predicates and 64-bit address arithmetic than on the
  // element, and keeps only what 80 registers hold in flight: 3.2 TB/s, profiles/r2_ncu_full_cfg5_s1_final.json.)
  if (fast) {
    const int rb0 = warp * (kFR / kWarps);
    double* o_p = out + (c0 + rb0) * p.ldo + i;
    const double* fu_p = p.in_unit.vec ? p.in_unit.vec + c0 + rb0 : nullptr;
    const double* gc_p = p.g_col + c0 + rb0;
    const int64_t diag = i - p.epi.diag_offset - c0 - rb0;             // result row (relative) that holds (r, r) for x = 0
    const uint32_t* t_p = tile + rb0 * kFPitch + 2 * lane;
    const double2* so_p = reinterpret_cast<const double2*>(so_s + rb0 * kFI) + lane;
    const uint32_t* cn_p = reinterpret_cast<const uint32_t*>(cn_s + rb0 * kFI) + lane;
    const bool addc = p.add_counts != 0, ev = p.use_evidence != 0;
#pragma unroll 4
    for (int rb = 0; rb < kFR / kWarps; ++rb) {
      const double2 so = so_p[rb * (kFI / 2)];
      const uint32_t cw = cn_p[rb * (kFI / 2)];
      const double fu = fu_p ? fu_p[rb] * p.in_unit.mul + p.in_unit.add : p.in_unit.add, gc = gc_p[rb];
      const uint32_t c0w = cw & 0xffffu, c1w = cw >> 16;
      double v0 = final_value_u16(g0, gc, u32_to_double(t_p[rb * kFPitch]), fu, addc ? u32_to_double(c0w) : 0.0, p.epi.coef,
                                  ev ? evidence_factor(c0w) : 1.0);
      double v1 = final_value_u16(g1, gc, u32_to_double(t_p[rb * kFPitch + 1]), fu, addc ? u32_to_double(c1w) : 0.0, p.epi.coef,
                                  ev ? evidence_factor(c1w) : 1.0);
      if (diag == rb) v0 = 1.0; else omax = fmax(omax, v0);
      if (diag + 1 == rb) v1 = 1.0; else omax = fmax(omax, v1);
      const double d0 = fabs(v0 - so.x), d1 = fabs(v1 - so.y);
      if (d0 > dmax) dmax = d0;                                        // NaN compares false: ignored like SimRank.py:74
      if (d1 > dmax) dmax = d1;
      __stcs(reinterpret_cast<double2*>(o_p + rb * p.ldo), make_double2(v0, v1));
    }
  } else
  for (int rb = warp * (kFR / kWarps); rb < (warp + 1) * (kFR / kWarps); rb += kFBatch) {
    double2 so[kFBatch];
    uint32_t c_lo[kFBatch], c_hi[kFBatch];                     // (16-bit counts: both in c_lo, split when used)
#pragma unroll
    for (int b = 0; b < kFBatch; ++b) {
      const int64_t r = c0 + rb + b;
      so[b] = make_double2(0.0, 0.0); c_lo[b] = c_hi[b] = 0u;
      if (r >= p.L || !in0) continue;
      if (p.epi.s_old) {
        const double* q = p.epi.s_old + r * p.epi.ld_s_old + i;
        if (vec) so[b] = __ldcs(reinterpret_cast<const double2*>(q));
        else { so[b].x = __ldcs(q); if (in1) so[b].y = __ldcs(q + 1); }
      }
      if (p.counts) {
        if (cvec) {
          c_lo[b] = __ldcs(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p.counts) + r * p.ld_counts + i));
        } else {
          c_lo[b] = load_count(p.counts, r * p.ld_counts + i, p.counts32);
          if (in1) c_hi[b] = load_count(p.counts, r * p.ld_counts + i + 1, p.counts32);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < kFBatch; ++b) {
      const int64_t r = c0 + rb + b;
      if (r >= p.L || !in0) continue;
      const double fu = row_bound(p.in_unit, r), gc = p.g_col[r];
      double v[2];
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (x && !in1) { v[1] = 0.0; continue; }
        const uint32_t cnt = cvec ? (x ? c_lo[b] >> 16 : c_lo[b] & 0xffffu) : (x ? c_hi[b] : c_lo[b]);
        double evf = 1.0;
        if (p.use_evidence) evf = evidence_factor(cnt);
        else if (p.epi.evidence) evf = evidence_factor(__ldcs(p.epi.evidence + r * p.epi.ld_evidence + i + x));
        double val = final_value_u16(x ? g1 : g0, gc, (double)tile[(rb + b) * kFPitch + 2 * lane + x], fu,
                                     p.add_counts ? (double)cnt : 0.0, p.epi.coef, evf);
        if (p.epi.prior) val = (1.0 - p.epi.lambda) * val + p.epi.lambda * __ldcs(p.epi.prior + r * p.epi.ld_prior + i + x);
        if (r + p.epi.diag_offset == i + x) val = 1.0; else if (val > omax) omax = val;
        if (p.epi.s_old) {
          const double d = fabs(val - (x ? so[b].y : so[b].x));
          if (d > dmax) dmax = d;                    // NaN compares false: ignored like SimRank.py:74
        }
        v[x] = val;
      }
      double* q = out + r * p.ldo + i;
      if (vec) __stcs(reinterpret_cast<double2*>(q), make_double2(v[0], v[1]));
      else { __stcs(q, v[0]); if (in1) __stcs(q + 1, v[1]); }
    }
  }
  dmax = warp_max(dmax);
  omax = warp_max(omax);
  if (lane == 0) { red[0][warp] = dmax; red[1][warp] = omax; }
  __syncthreads();
  if (warp == 0) {
    dmax = warp_max(lane < kWarps ? red[0][lane] : 0.0);
    omax = warp_max(lane < kWarps ? red[1][lane] : 0.0);
    if (lane == 0) {
      if (p.maxdiff && dmax > 0.0) atomic_max_nonneg(p.maxdiff, dmax);
      if (p.maxoff && omax > 0.0) atomic_max_nonneg(p.maxoff, omax);
    }
  }
}

// SRK_CSR_FINISH after a FIRST-type ACCUM: T = (A X)^T re-quantised, OUT[c, i] = rint(accum[i, c] * unit(c) *
// qmax / out_bound(i)) as uint16 -- the epilogue of MODE_FIRST as a streaming transposition (64 x 128 tile).
__global__ void __launch_bounds__(kThreads, 3)
csr_finish_first_kernel(const Params p) {
  __shared__ uint16_t tile[kFR * (kFI + 2)];
  constexpr int kP = kFI + 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i0 = p.row_begin + (int64_t)blockIdx.x * kFI, c0 = (int64_t)blockIdx.y * kFR;
  const int rows_here = (int)min((int64_t)kFI, p.row_end - i0);
  {
    const int64_t nxt = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + kPrefetchAhead;
    const int64_t ni = p.row_begin + (nxt % gridDim.x) * kFI, nc = (nxt / gridDim.x) * kFR;
    if (nc < p.L && !(p.flags & 8))
      prefetch_block(p.accum + ni * p.ld_accum + nc, p.ld_accum * 4, (int)min((int64_t)kFI, p.row_end - ni), kFR * 4);
  }
  {
    uint4 v[kFI / kWarps];
    double inv[kFI / kWarps];
#pragma unroll
    for (int q = 0; q < kFI / kWarps; ++q) {
      const int il = warp + kWarps * q;
      v[q] = make_uint4(0u, 0u, 0u, 0u);
      inv[q] = 0.0;
      if (il < rows_here) {
        v[q] = __ldcs(reinterpret_cast<const uint4*>(p.accum + (i0 + il) * p.ld_accum + c0 + 4 * lane));
        const double bo = row_bound(p.out_bound, i0 + il);
        inv[q] = bo > 0.0 ? p.qmax / bo : 0.0;
      }
    }
    double u[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) u[x] = c0 + 4 * lane + x < p.L ? row_bound(p.in_unit, c0 + 4 * lane + x) : 0.0;
#pragma unroll
    for (int q = 0; q < kFI / kWarps; ++q) {
      const int il = warp + kWarps * q;
      const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        // the expression of MODE_FIRST, rint(D unit / (bound / qmax)) clipped, without conversion instructions
        tile[(4 * lane + x) * kP + il] = (uint16_t)rint_clip_u32(u32_to_double(w[x]) * u[x] * inv[q], p.qmax);
      }
    }
  }
  __syncthreads();
  // out: a warp stores 16 rows of OUT, lane l the columns i0 + 2 l, i0 + 2 l + 1 (4 bytes per lane)
  const int64_t i = i0 + 2 * lane;
  uint16_t* out = reinterpret_cast<uint16_t*>(p.OUT);
  const bool pair = i + 1 < p.row_end && (p.ldo & 1) == 0 && (i & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  for (int rb = warp * (kFR / kWarps); rb < (warp + 1) * (kFR / kWarps); ++rb) {
    const int64_t r = c0 + rb;
    if (r >= p.L || i >= p.row_end) continue;
    const uint32_t lo = tile[rb * kP + 2 * lane], hi = tile[rb * kP + 2 * lane + 1];
    if (pair) *reinterpret_cast<uint32_t*>(out + r * p.ldo + i) = lo | (hi << 16);
    else { out[r * p.ldo + i] = (uint16_t)lo; if (i + 1 < p.row_end) out[r * p.ldo + i + 1] = (uint16_t)hi; }
  }
}

// SRK_CSR_FINISH of a SYMMETRIC second half: accum[i, r] holds D[i, r] for every r >= i (ACCUM with
// upper_only).  Tiles (64 rows i) x (64 columns r) with a column at or right of a row: inputs and the result row
// are read / written along r, the mirror image goes through a shared-memory transposition, so both copies
// leave as whole 512-byte runs (the fused symmetric launch scatters the mirror 8 bytes at a time).
constexpr int kSI = 64;
// shared memory: the transposition tile of the mirror store, then (fast path) the tiles of the sums, S_old and
// counts, staged by cp.async like in csr_finish_kernel
constexpr int kSTileBytes = kSI * (kSI + 1) * 8, kSSmemBytes = kSTileBytes + kSI * kSI * (4 + 8 + 2);
__global__ void __launch_bounds__(kThreads, 2)
csr_finish_sym_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t fin_smem[];
  double* tile = reinterpret_cast<double*>(fin_smem);
  uint32_t* acc_s = reinterpret_cast<uint32_t*>(fin_smem + kSTileBytes);
  double* so_s = reinterpret_cast<double*>(fin_smem + kSTileBytes + kSI * kSI * 4);
  uint16_t* cn_s = reinterpret_cast<uint16_t*>(fin_smem + kSTileBytes + kSI * kSI * 12);
  __shared__ double red[2][kWarps];
  constexpr int kP = kSI + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // triangular numbering of the block pairs (bi <= br)
  const int64_t t = blockIdx.x;
  int64_t br = (int64_t)((sqrt(1.0 + 8.0 * (double)t) - 1.0) * 0.5);
  while (br * (br + 1) / 2 > t) --br;
  while ((br + 1) * (br + 2) / 2 <= t) ++br;
  const int64_t bi = t - br * (br + 1) / 2;
  const int64_t i0 = bi * kSI, r0 = br * kSI;
  const int64_t r = r0 + 2 * lane;                                     // this lane's two columns
  const bool in0 = r < p.L, in1 = r + 1 < p.L;
  const bool vec = in1 && p.vec_aligned && ((p.ldo | p.epi.ld_s_old) & 1) == 0;
  const bool cvec = in1 && !p.counts32 && (p.ld_counts & 1) == 0 && (reinterpret_cast<uintptr_t>(p.counts) & 3) == 0;
  const double fu0 = in0 ? row_bound(p.in_unit, r) : 0.0, fu1 = in1 ? row_bound(p.in_unit, r + 1) : 0.0;
  const double gc0 = in0 ? p.g_col[r] : 0.0, gc1 = in1 ? p.g_col[r + 1] : 0.0;
  double* out = reinterpret_cast<double*>(p.OUT);
  double dmax = 0.0, omax = 0.0;
  constexpr int kRows = kSI / kWarps;                                  // 8 rows per warp, all loads first
  // Fast path, CTA-uniform: a full tile right of the diagonal blocks with the usual operands
  const bool fast = bi != br && r0 + kSI <= p.L && i0 + kSI <= p.row_end && p.vec_aligned && (p.ldo & 1) == 0 &&
                    (p.epi.ld_s_old & 1) == 0 && (p.ld_counts & 7) == 0 && p.epi.s_old && p.counts && !p.counts32 &&
                    !p.epi.evidence && !(p.flags & 16);
  if (fast) {
    const uint32_t* ab = p.accum + i0 * p.ld_accum + r0;
    const double* sb = p.epi.s_old + i0 * p.epi.ld_s_old + r0;
    const uint16_t* cb = reinterpret_cast<const uint16_t*>(p.counts) + i0 * p.ld_counts + r0;
#pragma unroll
    for (int k = 0; k < kSI * kSI * 4 / 16 / kThreads; ++k) {
      const int ch = threadIdx.x + kThreads * k, row = ch >> 4, c16 = ch & 15;
      cp_async16(acc_s + row * kSI + 4 * c16, ab + row * p.ld_accum + 4 * c16);
    }
#pragma unroll
    for (int k = 0; k < kSI * kSI * 8 / 16 / kThreads; ++k) {
      const int ch = threadIdx.x + kThreads * k, row = ch >> 5, c16 = ch & 31;
      cp_async16(so_s + row * kSI + 2 * c16, sb + row * p.epi.ld_s_old + 2 * c16);
    }
#pragma unroll
    for (int k = 0; k < kSI * kSI * 2 / 16 / kThreads; ++k) {
      const int ch = threadIdx.x + kThreads * k, row = ch >> 3, c16 = ch & 7;
      cp_async16(cn_s + row * kSI + 8 * c16, cb + row * p.ld_counts + 8 * c16);
    }
    cp_async_wait_all();
    __syncthreads();
    const bool addc = p.add_counts != 0, ev = p.use_evidence != 0;
#pragma unroll 4
    for (int q = 0; q < kRows; ++q) {
      const int il = warp * kRows + q;
      const int64_t i = i0 + il;
      const uint2 sum = *reinterpret_cast<const uint2*>(acc_s + il * kSI + 2 * lane);
      const double2 so = *reinterpret_cast<const double2*>(so_s + il * kSI + 2 * lane);
      const uint32_t cw = *reinterpret_cast<const uint32_t*>(cn_s + il * kSI + 2 * lane);
      const uint32_t c0w = cw & 0xffffu, c1w = cw >> 16;
      const double gi = p.g[i];
      const double v0 = final_value_u16(gi, gc0, u32_to_double(sum.x), fu0, addc ? u32_to_double(c0w) : 0.0, p.epi.coef,
                                        ev ? evidence_factor(c0w) : 1.0);
      const double v1 = final_value_u16(gi, gc1, u32_to_double(sum.y), fu1, addc ? u32_to_double(c1w) : 0.0, p.epi.coef,
                                        ev ? evidence_factor(c1w) : 1.0);
      omax = fmax(omax, fmax(v0, v1));
      const double d0 = fabs(v0 - so.x), d1 = fabs(v1 - so.y);
      if (d0 > dmax) dmax = d0;                                        // NaN compares false: ignored like SimRank.py:74
      if (d1 > dmax) dmax = d1;
      __stcs(reinterpret_cast<double2*>(out + i * p.ldo + r), make_double2(v0, v1));
      tile[(2 * lane) * kP + il] = v0;
      tile[(2 * lane + 1) * kP + il] = v1;
    }
  } else {
  uint2 sum[kRows];
  double2 so[kRows];
  uint32_t cw[kRows], ch[kRows];
#pragma unroll
  for (int q = 0; q < kRows; ++q) {
    const int64_t i = i0 + warp * kRows + q;
    sum[q] = make_uint2(0u, 0u); so[q] = make_double2(0.0, 0.0); cw[q] = ch[q] = 0u;
    if (i >= p.row_end || !in0) continue;
    sum[q] = __ldcs(reinterpret_cast<const uint2*>(p.accum + i * p.ld_accum + r));
    if (p.epi.s_old) {
      const double* s = p.epi.s_old + i * p.epi.ld_s_old + r;
      if (vec) so[q] = __ldcs(reinterpret_cast<const double2*>(s));
      else { so[q].x = __ldcs(s); if (in1) so[q].y = __ldcs(s + 1); }
    }
    if (p.counts) {
      if (cvec) cw[q] = __ldcs(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p.counts) + i * p.ld_counts + r));
      else { cw[q] = load_count(p.counts, i * p.ld_counts + r, p.counts32); if (in1) ch[q] = load_count(p.counts, i * p.ld_counts + r + 1, p.counts32); }
    }
  }
#pragma unroll
  for (int q = 0; q < kRows; ++q) {
    const int il = warp * kRows + q;
    const int64_t i = i0 + il;
    double v[2] = {0.0, 0.0};
    if (i < p.row_end && in0) {
      const double gi = p.g[i];
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int64_t c = r + x;
        if ((x && !in1) || c < i) continue;                            // left of the diagonal: the mirror of (c, i)
        const uint32_t cnt = cvec ? (x ? cw[q] >> 16 : cw[q] & 0xffffu) : (x ? ch[q] : cw[q]);
        double evf = 1.0;
        if (p.use_evidence) evf = evidence_factor(cnt);
        else if (p.epi.evidence) evf = evidence_factor(__ldcs(p.epi.evidence + i * p.epi.ld_evidence + c));
        double val = final_value_u16(gi, x ? gc1 : gc0, (double)(x ? sum[q].y : sum[q].x), x ? fu1 : fu0,
                                     p.add_counts ? (double)cnt : 0.0, p.epi.coef, evf);
        if (c == i) val = 1.0; else if (val > omax) omax = val;
        if (p.epi.s_old) {
          const double d = fabs(val - (x ? so[q].y : so[q].x));
          if (d > dmax) dmax = d;                                      // NaN compares false: ignored like SimRank.py:74
        }
        v[x] = val;
      }
      // the row itself: both columns at or right of the diagonal -> one 16-byte store
      double* o = out + i * p.ldo + r;
      if (vec && r >= i) __stcs(reinterpret_cast<double2*>(o), make_double2(v[0], v[1]));
      else { if (r >= i) __stcs(o, v[0]); if (in1 && r + 1 >= i) __stcs(o + 1, v[1]); }
    }
    tile[(2 * lane) * kP + il] = v[0];
    tile[(2 * lane + 1) * kP + il] = v[1];
  }
  }
  __syncthreads();
  // mirror: element (c, i) for c > i; a warp writes 8 rows c of the result, lane l the columns i0 + 2 l, + 1
  {
    const int64_t i = i0 + 2 * lane;
#pragma unroll
    for (int q = 0; q < kRows; ++q) {
      const int cl = warp * kRows + q;
      const int64_t c = r0 + cl;
      if (c >= p.L || i >= p.row_end) continue;
      const double a = tile[cl * kP + 2 * lane], b = tile[cl * kP + 2 * lane + 1];
      double* o = out + c * p.ldo + i;
      const bool two = i + 1 < p.row_end;
      if (two && c > i + 1 && p.vec_aligned && (p.ldo & 1) == 0) __stcs(reinterpret_cast<double2*>(o), make_double2(a, b));
      else { if (c > i) __stcs(o, a); if (two && c > i + 1) __stcs(o + 1, b); }
    }
  }
  dmax = warp_max(dmax);
  omax = warp_max(omax);
  if (lane == 0) { red[0][warp] = dmax; red[1][warp] = omax; }
  __syncthreads();
  if (warp == 0) {
    dmax = warp_max(lane < kWarps ? red[0][lane] : 0.0);
    omax = warp_max(lane < kWarps ? red[1][lane] : 0.0);
    if (lane == 0) {
      if (p.maxdiff && dmax > 0.0) atomic_max_nonneg(p.maxdiff, dmax);
      if (p.maxoff && omax > 0.0) atomic_max_nonneg(p.maxoff, omax);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fixed-point source operand of the u16 gather.  unit[r] = max_k V[r, k] / qmax over the row (the
// element (r, r + zero_diag_offset) excluded: S = I + S_off), then XT[k, r] = rint(V[r, k] * (1 / unit[r])):
// the TRANSPOSED matrix with one scale per column -- for a symmetric S this is S_off itself scaled by
// its column maxima, which is what lets the gather sum whole columns as integers.
constexpr int QT = 64;
__global__ void __launch_bounds__(256)
row_unit_kernel(const double* __restrict__ V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                double qmax, double* __restrict__ unit) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = blockIdx.x;
  const double* row = V + r * ldv;
  const int64_t kd = zero_diag_offset >= 0 ? r + zero_diag_offset : -1;
  double m = 0.0;
  for (int64_t k = threadIdx.x; k < K; k += 256) {
    const double v = row[k];
    if (k != kd && v > m) m = v;                               // NaN and negatives never win
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (warp == 0) {
    m = lane < 8 ? red[lane] : 0.0;
    m = warp_max(m);
    if (lane == 0) unit[r] = m / qmax;
  }
}
__global__ void __launch_bounds__(256)
quantize_transpose_u16_kernel(const double* __restrict__ V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                              double qmax, const double* __restrict__ unit, uint16_t* __restrict__ XT, int64_t ldxt) {
  __shared__ uint16_t t[QT][QT + 2];
  __shared__ double inv[QT];
  const int64_t r0 = (int64_t)blockIdx.y * QT, k0 = (int64_t)blockIdx.x * QT;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
  if (threadIdx.x < QT) {
    const double u = r0 + threadIdx.x < R ? unit[r0 + threadIdx.x] : 0.0;
    inv[threadIdx.x] = u > 0.0 ? 1.0 / u : 0.0;                // one division per row of the tile, not per element
  }
  __syncthreads();
  for (int rr = ty; rr < QT; rr += 4) {
    const int64_t r = r0 + rr, k = k0 + tx;
    unsigned q = 0u;
    if (r < R && k < K) {
      const double v = (zero_diag_offset >= 0 && k == r + zero_diag_offset) ? 0.0 : V[r * ldv + k];
      q = rint_clip_u32(v * inv[rr], qmax);
    }
    t[tx][rr] = (uint16_t)q;
  }
  __syncthreads();
  for (int kk = ty; kk < QT; kk += 4) {
    const int64_t k = k0 + kk, r = r0 + tx;
    if (k < K && r < ldxt) XT[k * ldxt + r] = r < R ? t[kk][tx] : (uint16_t)0;
  }
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

// Symmetric V (R == K, bit-exactly symmetric as the symmetric second half leaves it): XT[k, r] =
// rint(V[r, k] / unit[r]) = rint(V[k, r] / unit[r]) -- a streaming pass, no transposition.
__global__ void __launch_bounds__(256)
quantize_sym_u16_kernel(const double* __restrict__ V, int64_t ldv, int64_t n, int64_t zero_diag_offset, double qmax,
                        const double* __restrict__ unit, uint16_t* __restrict__ XT, int64_t ldxt) {
  // A warp covers 8 rows x 64 consecutive columns: lane l takes the column pair 2 l of every row (512 B
  // per load instruction, 128 B per store instruction, eight independent loads in flight) and divides
  // by the units of its two columns ONCE: values are scaled with the reciprocal (a float64 division per
  // element made the pass compute-bound).
  const int lane = threadIdx.x & 31;
  const int64_t chunks = (ldxt + 63) / 64, row_blocks = (n + 7) / 8;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= row_blocks * chunks) return;
  const int64_t k0 = (w / chunks) * 8, r = (w % chunks) * 64 + 2 * lane;
  if (r >= ldxt) return;
  const bool vec = ((reinterpret_cast<uintptr_t>(V) | (uintptr_t)(ldv * 8)) & 15) == 0;
  double inv[2];
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const double u = r + x < n ? unit[r + x] : 0.0;
    inv[x] = u > 0.0 ? 1.0 / u : 0.0;
  }
  double2 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t k = k0 + j;
    if (k >= n) v[j] = make_double2(0.0, 0.0);
    else if (vec && r + 1 < n) v[j] = __ldcs(reinterpret_cast<const double2*>(V + k * ldv + r));
    else v[j] = make_double2(r < n ? V[k * ldv + r] : 0.0, r + 1 < n ? V[k * ldv + r + 1] : 0.0);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t k = k0 + j;
    if (k >= n) continue;
    uint32_t out = 0u;
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const double val = (zero_diag_offset >= 0 && k == r + x + zero_diag_offset) ? 0.0 : (x ? v[j].y : v[j].x);
      out |= rint_clip_u32(val * inv[x], qmax) << (16 * x);
    }
    *reinterpret_cast<uint32_t*>(XT + k * ldxt + r) = out;
  }
}

template <typename E, int TC, int MODE>
static int launch_one(Params& p, int64_t x_rows, dim3 grid, cudaStream_t st) {
  typedef Smem<E, TC, MODE> SM;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (p.tma) {
    // X as a 2-D tensor of 4-byte elements [x_rows][ldx * sizeof(E) / 4]; a box is one row segment
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(SRK_ERR_CUDA, "%s", "cuTensorMapEncodeTiled is not available from the driver");
    const int64_t row_bytes = p.ldx * (int64_t)sizeof(E);
    cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / 4), (cuuint64_t)x_rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(SM::kSeg / 4), 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(p.X), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) p.tma = 0;                                   // shapes TMA cannot describe: plain loads
  }
  auto kern = csr_gather_kernel<E, TC, MODE>;
  SRK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
  kern<<<grid, kThreads, SM::kBytes, st>>>(map, p);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

}  // namespace gat
}  // namespace srk

using namespace srk;

extern "C" int srk_csr_half(const srk_csr_args* a, void* stream) {
  SRK_REQUIRE(a, "null args");
  const bool accum_mode = a->mode == SRK_CSR_ACCUM, finish_first = a->mode == SRK_CSR_FINISH_FIRST;
  const bool finish_mode = a->mode == SRK_CSR_FINISH || finish_first;
  SRK_REQUIRE(finish_mode || (a->indices && a->X), "null pointer");
  SRK_REQUIRE(accum_mode || (a->g && a->OUT && (finish_mode || a->indptr)), "null pointer");
  SRK_REQUIRE((a->row_lo == nullptr) == (a->row_hi == nullptr), "row_lo and row_hi come together");
  SRK_REQUIRE(0 <= a->row_begin && a->row_begin <= a->row_end && a->row_end <= a->M, "row range");
  // OUT is addressed as OUT[c * ldo + i] for i in [row_begin, row_end) only: a caller that stores just
  // those columns passes the address of (virtual) column 0, i.e. its buffer minus row_begin elements
  SRK_REQUIRE(a->L >= 0 && (finish_mode || a->ldx >= a->L) && (accum_mode || a->ldo >= a->row_end - a->row_begin),
              "leading dimensions");
  SRK_REQUIRE(a->K >= 0 && a->K < (1ll << 31), "K (rows of X) out of range");
  SRK_REQUIRE(a->elem == SRK_ELEM_F64 || a->elem == SRK_ELEM_U16, "elem must be SRK_ELEM_F64 or SRK_ELEM_U16");
  SRK_REQUIRE(a->mode == SRK_CSR_FIRST || a->mode == SRK_CSR_FINAL || accum_mode || finish_mode,
              "mode must be SRK_CSR_FIRST, SRK_CSR_FINAL, SRK_CSR_ACCUM, SRK_CSR_FINISH or SRK_CSR_FINISH_FIRST");
  if (accum_mode) SRK_REQUIRE(a->row_lo && a->accum && a->accum_slot, "SRK_CSR_ACCUM needs row_lo, row_hi, accum and accum_slot");
  if (finish_mode) SRK_REQUIRE(a->accum, "SRK_CSR_FINISH needs accum (one row of sums per graph row)");
  if (a->accum || a->accum_slot) {
    SRK_REQUIRE(a->elem == SRK_ELEM_U16, "pre-summed pieces exist in the fixed-point mode only");
    SRK_REQUIRE(a->accum && (a->accum_slot || finish_mode) && a->ld_accum % 512 == 0 && a->ld_accum >= a->L &&
                    ((uintptr_t)a->accum % 32) == 0,
                "accum: 32-byte aligned, ld_accum a multiple of 512 and >= L");
  }
  SRK_REQUIRE(a->counts_bits == 0 || a->counts_bits == 16 || a->counts_bits == 32, "counts_bits must be 16 or 32");
  SRK_REQUIRE(!(a->add_counts || a->use_evidence) || a->counts, "counts missing");
  SRK_REQUIRE(!(a->use_evidence && a->epi.evidence), "evidence given twice (counts and epi.evidence)");
  const bool sym = (a->mode == SRK_CSR_FINAL || a->mode == SRK_CSR_FINISH) && a->symmetric;
  if (sym)
    SRK_REQUIRE(a->row_begin == 0 && a->row_end == a->M && a->L == a->M && a->epi.prior == nullptr &&
                    a->epi.diag_offset == 0 && a->ldo >= a->L,
                "the symmetric second half needs the whole square problem and no prior");
  if (a->elem == SRK_ELEM_U16 && (a->mode == SRK_CSR_FINAL || a->mode == SRK_CSR_FINISH)) SRK_REQUIRE(a->g_col, "u16 FINAL needs g_col");
  if (a->row_end == a->row_begin || a->L == 0) return SRK_OK;

  gat::Params p;
  memset(&p, 0, sizeof(p));
  p.rowbeg = a->row_lo ? a->row_lo : a->indptr;
  p.rowend = a->row_hi ? a->row_hi : a->indptr + 1;
  p.indices = a->indices; p.g = a->g;
  p.accum = a->accum; p.ld_accum = a->ld_accum; p.accum_slot = a->accum_slot;
  p.row_begin = a->row_begin; p.row_end = a->row_end;
  p.X = a->X; p.ldx = a->ldx; p.L = a->L; p.OUT = a->OUT; p.ldo = a->ldo;
  p.in_unit = a->in_unit; p.out_bound = a->out_bound; p.g_col = a->g_col;
  p.qmax = a->qmax == 0.0 ? 65535.0 : a->qmax;
  SRK_REQUIRE(p.qmax >= 1.0 && p.qmax <= 65535.0, "qmax must be in 1..65535");
  p.counts = a->counts; p.ld_counts = a->ld_counts; p.counts32 = a->counts_bits == 32;
  p.add_counts = a->add_counts; p.use_evidence = a->use_evidence;
  p.upper_only = accum_mode && a->symmetric;
  if (a->mode == SRK_CSR_FINAL || a->mode == SRK_CSR_FINISH) { p.epi = to_dev(a->epi); p.maxdiff = a->epi.maxdiff; p.maxoff = a->epi.maxoff; }
  const int64_t esz = a->elem == SRK_ELEM_U16 ? 2 : 8;
  // TMA needs 16-byte aligned rows
  p.tma = ((uintptr_t)a->X % 16 == 0) && ((a->ldx * esz) % 16 == 0);
  p.vec_aligned = (((uintptr_t)a->OUT | (uintptr_t)a->epi.s_old | (uintptr_t)a->counts) % 16) == 0;
  const int64_t x_rows = a->K > 0 ? a->K : (1ll << 31) - 1;        // bound of TMA's row check (K unknown: none)
  { const char* e = getenv("SRK_CSR_FLAGS"); p.flags = e ? atoi(e) : 0; if (p.flags & 2) p.tma = 0; }

  // Panel width = columns of X per CTA = what all CTAs of a grid column gather from; the panel (rows
  // of X x segment bytes, held once per L2 die) has to survive in L2 next to the streams of the
  // epilogue.  1 KB segments (uint16: 512 columns, float64: 128) reach the L2 roof, 512 B ones stop at
  // two thirds of it (profiles/r2_micro_tma_gather_rate.txt); the transposed second half keeps a 4- or
  // 8-byte tile per element in shared memory and takes 16 graph rows per CTA instead of 32.
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = a->row_end - a->row_begin;
  if (finish_mode && sym) {
    const int64_t nb = (a->M + gat::kSI - 1) / gat::kSI, total = nb * (nb + 1) / 2;
    SRK_REQUIRE(total < (1ll << 31), "too many tiles");
    SRK_CUDA_OK(cudaFuncSetAttribute(gat::csr_finish_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gat::kSSmemBytes));
    gat::csr_finish_sym_kernel<<<(unsigned)total, gat::kThreads, gat::kSSmemBytes, st>>>(p);
    SRK_CUDA_OK(cudaGetLastError());
    return SRK_OK;
  }
  if (finish_mode) {
    const int64_t fx = (rows + gat::kFI - 1) / gat::kFI, fy = (a->L + gat::kFR - 1) / gat::kFR;
    SRK_REQUIRE(fy <= 65535, "too many column panels");
    if (finish_first) {
      gat::csr_finish_first_kernel<<<dim3((unsigned)fx, (unsigned)fy), gat::kThreads, 0, st>>>(p);
      SRK_CUDA_OK(cudaGetLastError());
      return SRK_OK;
    }
    // two CTAs per SM: the S_old / counts tiles of the fast path are staged in shared memory (113 KB per CTA);
    // three (80 registers) and four CTAs (64, spills) were tried without staging: profiles/r2_csr_shapes_finish_ctas.jsonl
    auto kern = gat::csr_finish_kernel<2>;
    SRK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gat::kFSmemBytes));
    kern<<<dim3((unsigned)fx, (unsigned)fy), gat::kThreads, gat::kFSmemBytes, st>>>(p);
    SRK_CUDA_OK(cudaGetLastError());
    return SRK_OK;
  }
  const int mode = accum_mode ? gat::MODE_ACCUM
                              : a->mode == SRK_CSR_FIRST ? gat::MODE_FIRST : (sym ? gat::MODE_FINAL_SYM : gat::MODE_FINAL);
  // (Narrower panels were tried for operands with so many rows that a 1 KB-wide panel cannot stay in L2
  // -- 138k rows at BASELINE cfg5: 142 MB -- and did not pay: 512 B segments cost more per byte than the
  // residency wins back, profiles/r2_csr_shapes.jsonl.)
  const int tc = a->elem == SRK_ELEM_U16 ? 512 : 128;
  const int ti = mode == gat::MODE_FINAL ? 16 : (accum_mode ? gat::kTIAccum : gat::kTI);
  const int64_t gx = (rows + ti - 1) / ti, gy = (a->L + tc - 1) / tc;
  SRK_REQUIRE(gy <= 65535, "too many column panels");
  dim3 grid((unsigned)gx, (unsigned)gy);
  p.tiles_x = gx;
  if (sym) {
    const int64_t per = tc / ti, full = gx / per;                    // see the kernel: triangular numbering
    int64_t total = per * full * (full + 1) / 2;
    if (gy > full) total += (gy - full) * gx;
    SRK_REQUIRE(total < (1ll << 31), "too many tiles");
    grid = dim3((unsigned)total, 1);
  }
  if (a->elem == SRK_ELEM_U16) {
    switch (mode) {
      case gat::MODE_FIRST: return gat::launch_one<uint16_t, 512, gat::MODE_FIRST>(p, x_rows, grid, st);
      case gat::MODE_FINAL: return gat::launch_one<uint16_t, 512, gat::MODE_FINAL>(p, x_rows, grid, st);
      case gat::MODE_ACCUM: return gat::launch_one<uint16_t, 512, gat::MODE_ACCUM>(p, x_rows, grid, st);
      default: return gat::launch_one<uint16_t, 512, gat::MODE_FINAL_SYM>(p, x_rows, grid, st);
    }
  }
  switch (mode) {
    case gat::MODE_FIRST: return gat::launch_one<double, 128, gat::MODE_FIRST>(p, x_rows, grid, st);
    case gat::MODE_FINAL: return gat::launch_one<double, 128, gat::MODE_FINAL>(p, x_rows, grid, st);
    default: return gat::launch_one<double, 128, gat::MODE_FINAL_SYM>(p, x_rows, grid, st);
  }
}

extern "C" int srk_csr_half_f64(const int64_t* indptr, const int32_t* indices, const double* g,
                                int64_t M, int64_t row_begin, int64_t row_end, const double* X,
                                int64_t ldx, int64_t L, double* OUT, int64_t ldo,
                                const srk_epilogue* final_epi, void* stream) {
  srk_csr_args a;
  memset(&a, 0, sizeof(a));
  a.elem = SRK_ELEM_F64;
  a.mode = final_epi ? SRK_CSR_FINAL : SRK_CSR_FIRST;
  a.indptr = indptr; a.indices = indices; a.g = g;
  a.M = M; a.row_begin = row_begin; a.row_end = row_end;
  a.X = X; a.ldx = ldx; a.L = L; a.OUT = OUT; a.ldo = ldo;
  if (final_epi) a.epi = *final_epi;
  return srk_csr_half(&a, stream);
}

extern "C" int srk_quantize_rows_u16(const double* V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                                     uint16_t* XT, int64_t ldxt, double* unit, double qmax, int symmetric,
                                     void* stream) {
  SRK_REQUIRE(V && XT && unit, "null pointer");
  if (qmax == 0.0) qmax = 65535.0;
  SRK_REQUIRE(qmax >= 1.0 && qmax <= 65535.0 && qmax == floor(qmax), "qmax must be an integer in 1..65535");
  SRK_REQUIRE(ldv >= K && ldxt >= R && ldxt % 8 == 0 && ((uintptr_t)XT % 16) == 0,
              "XT must be 16-byte aligned with ldxt a multiple of 8 and >= R");
  SRK_REQUIRE(!symmetric || R == K, "a symmetric matrix is square");
  if (R == 0 || K == 0) return SRK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  gat::row_unit_kernel<<<(unsigned)R, 256, 0, st>>>(V, ldv, R, K, zero_diag_offset, qmax, unit);
  if (symmetric) {
    const int64_t threads = ((K + 7) / 8) * ((ldxt + 63) / 64) * 32;
    SRK_REQUIRE((threads + 255) / 256 < (1ll << 31), "matrix too large for one launch");
    gat::quantize_sym_u16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(V, ldv, K, zero_diag_offset, qmax, unit, XT, ldxt);
    SRK_CUDA_OK(cudaGetLastError());
    return SRK_OK;
  }
  dim3 grid((unsigned)((K + gat::QT - 1) / gat::QT), (unsigned)((ldxt + gat::QT - 1) / gat::QT));
  SRK_REQUIRE(grid.y <= 65535, "too many row tiles");
  gat::quantize_transpose_u16_kernel<<<grid, 256, 0, st>>>(V, ldv, R, K, zero_diag_offset, qmax, unit, XT, ldxt);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}
