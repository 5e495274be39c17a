// CSR half-products: the neighbour rows of X are gathered through shared memory by bulk
// asynchronous copies (cp.async.bulk, the 1-D form of TMA) and summed by the warp that owns the
// graph row.  Two arithmetic modes share the kernel:
//
//   f64  OUT[c, i] = g[i] * sum_{m in N(i)} X[m, c]                      exact float64 (DADD)
//   u16  D[i, c]   = sum_{m in N(i)} Xq[m, c]    Xq uint16 fixed point with one scale per COLUMN of
//        X, so the sum over m is an exact integer (deg < 65536 keeps it inside 32 bits); the scales
//        are applied once per output element in the epilogue.  4x fewer gathered bytes.
//
// Why it looks like this (DESIGN.md "K3").  The streamed operands are O(n^2) bytes, the gather is
// nnz * n * sizeof(element): 550 GB per half-product at BASELINE cfg4 in float64, served by L2 (all
// CTAs of a grid column share one column panel of X, which is what keeps it there).  Measured on
// the previous kernel (ncu, profiles/r2_ncu_full_csr_f64_baseline.json): 549 GB over the
// L2->SM crossbar in 35.4 ms = 15.5 TB/s, L1 hit rate 0 -- the half-product sits on the L2
// bandwidth roof, not on HBM (27 GB of DRAM traffic).  So the bytes are cut (uint16 planes of the
// same row-max-scaled fixed point the tensor-core path uses, symmetric second half) and the loads
// are taken off the register file: every warp keeps kDepth row segments in flight in a private
// shared-memory ring, one elected lane per segment issues the copy, completion is an mbarrier
// transaction count, and the lanes read the landed segment with conflict-free ld.shared.
//
// CTA tile: TI = 32 graph rows x TC columns of X (TC * sizeof(element) = 512 B or 1 KB segments).
// Warp w owns the 4 CONSECUTIVE graph rows i0 + 4w .. i0 + 4w + 3: their neighbour lists are one
// contiguous range of `indices`, which the warp walks as a single stream -- the ring never drains
// at a row boundary.
//
// Epilogues
//   FIRST            T = (G X)^T: per finished row the values go to a shared-memory tile, the CTA
//                    writes the tile transposed (rows of OUT are contiguous in i).  u16: re-quantised
//                    with the bound of output column i (a per-node vector times a scalar).
//   FINAL            the same transposed store with the fused SimRank epilogue (srk_epilogue).
//   FINAL symmetric  square problems whose result is symmetric (no prior): only tiles that contain
//                    an element c >= i are computed, the epilogue runs in the row-major orientation
//                    straight from the accumulators (coalesced loads of S_old / counts, coalesced
//                    store of row i) and every off-diagonal value is ALSO stored at (c, i): each
//                    unordered pair is computed once, S stays bit-exactly symmetric, the second
//                    half gathers 3/4 .. 1/2 of the bytes.
#include <stdlib.h>

#include "common.cuh"

namespace srk {
namespace gat {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int TI = 32;                      // graph rows per CTA
constexpr int kRowsPerWarp = TI / kWarps;   // consecutive rows per warp
constexpr int kDepth = 8;                   // row segments in flight per warp (power of two, <= 32)

constexpr int MODE_FIRST = 0, MODE_FINAL = 1, MODE_FINAL_SYM = 2;

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
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
// global -> this CTA's shared memory, `bytes` (multiple of 16) counted on `bar`
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
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

struct Params {
  const int64_t* indptr; const int32_t* indices; const double* g;
  int64_t row_begin, row_end;
  const void* X; int64_t ldx, L;
  void* OUT; int64_t ldo;
  srk_rowbound in_unit;        // u16: value of one unit of column c of X
  srk_rowbound out_bound;      // u16 FIRST: bound of output column i (one unit = bound / 65535)
  const double* g_col;         // u16 FINAL: row factor of output row r (= column of X)
  const void* counts; int64_t ld_counts; int counts32, add_counts, use_evidence;
  EpilogueDev epi; double* maxdiff; double* maxoff;
  int bulk;                    // rows of X are 16-byte aligned: bulk copies; else plain loads
  int flags;                   // SRK_CSR_FLAGS (A/B profiling): 1 = default L2 policy for the gather
};

__device__ __forceinline__ uint32_t load_count(const void* base, int64_t idx, int c32) {
  return c32 ? reinterpret_cast<const uint32_t*>(base)[idx] : (uint32_t)reinterpret_cast<const uint16_t*>(base)[idx];
}

// Accumulators of one graph row: TC columns spread over the 32 lanes.
template <typename E, int TC>
struct Acc;

// float64: lane holds columns lane + 32 k
template <int TC>
struct Acc<double, TC> {
  static constexpr int kCols = TC / 32;
  double v[kCols];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < kCols; ++k) v[k] = 0.0;
  }
  __device__ __forceinline__ static int col(int j, int lane) { return lane + 32 * j; }
  __device__ __forceinline__ void add_smem(const uint8_t* seg, int lane) {
    const double* s = reinterpret_cast<const double*>(seg);
#pragma unroll
    for (int k = 0; k < kCols; ++k) v[k] += s[lane + 32 * k];
  }
  __device__ __forceinline__ void add_global(const void* row, int lane, int64_t valid) {
    const double* s = reinterpret_cast<const double*>(row);
#pragma unroll
    for (int k = 0; k < kCols; ++k) v[k] += (lane + 32 * k < valid) ? __ldg(s + lane + 32 * k) : 0.0;
  }
  __device__ __forceinline__ double val(int j) const { return v[j]; }
};

// uint16: lane holds the 32-bit words lane + 32 w, i.e. columns 2 (lane + 32 w) and + 1.  The low
// halves are not masked out per element: aw accumulates the whole words modulo 2^32 and ah the high
// halves, so sum(lo) = aw - (ah << 16) (mod 2^32), which is exact because sum(lo) < 2^32.
template <int TC>
struct Acc<uint16_t, TC> {
  static constexpr int kWords = TC / 64;
  static constexpr int kCols = 2 * kWords;
  uint32_t aw[kWords], ah[kWords];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int w = 0; w < kWords; ++w) aw[w] = ah[w] = 0u;
  }
  __device__ __forceinline__ static int col(int j, int lane) { return 2 * (lane + 32 * (j >> 1)) + (j & 1); }
  __device__ __forceinline__ void add_smem(const uint8_t* seg, int lane) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(seg);
#pragma unroll
    for (int w = 0; w < kWords; ++w) {
      const uint32_t x = s[lane + 32 * w];
      aw[w] += x;
      ah[w] += x >> 16;
    }
  }
  __device__ __forceinline__ void add_global(const void* row, int lane, int64_t valid) {
    const uint16_t* s = reinterpret_cast<const uint16_t*>(row);
#pragma unroll
    for (int w = 0; w < kWords; ++w) {
      const int c = 2 * (lane + 32 * w);
      const uint32_t lo = c < valid ? (uint32_t)__ldg(s + c) : 0u, hi = c + 1 < valid ? (uint32_t)__ldg(s + c + 1) : 0u;
      aw[w] += lo | (hi << 16);
      ah[w] += hi;
    }
  }
  __device__ __forceinline__ uint32_t raw(int j) const { return (j & 1) ? ah[j >> 1] : aw[j >> 1] - (ah[j >> 1] << 16); }
  __device__ __forceinline__ double val(int j) const { return (double)raw(j); }
};

template <typename E, int MODE>
struct TileElem { typedef double type; };
template <>
struct TileElem<uint16_t, MODE_FIRST> { typedef uint16_t type; };
template <>
struct TileElem<uint16_t, MODE_FINAL> { typedef uint32_t type; };

template <typename E, int TC, int MODE>
struct Smem {
  typedef typename TileElem<E, MODE>::type TileT;
  static constexpr int kSeg = TC * (int)sizeof(E);
  static constexpr int kRing = kWarps * kDepth * kSeg;
  static constexpr int kBars = kWarps * kDepth * 8;
  // transposed-store tile [TC][kPitch]: odd pitch in 32-bit words where the element size allows
  static constexpr int kPitch = sizeof(TileT) == 2 ? TI + 2 : TI + 1;
  static constexpr int kTile = MODE == MODE_FINAL_SYM ? 0 : TC * kPitch * (int)sizeof(TileT);
  static constexpr int kBytes = kRing + kBars + kTile;
};

template <typename E, int TC, int MODE>
__global__ void __launch_bounds__(kThreads)
csr_gather_kernel(const Params p) {
  typedef Smem<E, TC, MODE> SM;
  typedef typename SM::TileT TileT;
  constexpr bool kU16 = sizeof(E) == 2;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ double red[2][kWarps];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kRing);
  TileT* tile = reinterpret_cast<TileT*>(smem + SM::kRing + SM::kBars);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i0 = p.row_begin + (int64_t)blockIdx.x * TI;
  const int64_t c0 = (int64_t)blockIdx.y * TC;
  if (MODE == MODE_FINAL_SYM && c0 + TC - 1 < i0) return;      // every element has c < i: mirrored from above

  uint8_t* wring = smem + (size_t)warp * kDepth * SM::kSeg;
  uint64_t* wbar = bars + warp * kDepth;
  if (p.bulk) {
    if (lane < kDepth) mbar_init(&wbar[lane], 1);
    fence_barrier_init();
    __syncwarp();
  }

  double dmax = 0.0, omax = 0.0;
  const int64_t r_lo = i0 + (int64_t)warp * kRowsPerWarp;
  const int64_t r_hi = min(r_lo + (int64_t)kRowsPerWarp, p.row_end);
  if (r_lo < p.row_end) {
    const int64_t eb = p.indptr[r_lo], ee = p.indptr[r_hi];
    const int64_t valid = min((int64_t)TC, p.ldx - c0);               // columns of X this panel can read
    const uint32_t seg_bytes = (uint32_t)(valid * (int64_t)sizeof(E));
    const uint8_t* xbase = reinterpret_cast<const uint8_t*>(p.X) + c0 * (int64_t)sizeof(E);
    const int64_t pitch = p.ldx * (int64_t)sizeof(E);
    const uint64_t pol = (p.flags & 1) ? policy_evict_normal() : policy_evict_last();
    int idx_cur = (eb + lane < ee) ? p.indices[eb + lane] : 0;
    int idx_nxt = (eb + 32 + lane < ee) ? p.indices[eb + 32 + lane] : 0;
    int64_t chunk_base = eb;                                        // stream position held by lane 0 of idx_cur
    if (p.bulk && lane < kDepth && eb + lane < ee) {
      mbar_expect_tx(&wbar[lane], seg_bytes);
      bulk_copy(wring + lane * SM::kSeg, xbase + (int64_t)idx_cur * pitch, seg_bytes, &wbar[lane], pol);
    }
    Acc<E, TC> acc;
    acc.clear();
    int64_t e = eb;
    for (int64_t row = r_lo; row < r_hi; ++row) {
      const int64_t rend = p.indptr[row + 1];
      for (; e < rend; ++e) {
        if (e - chunk_base == 32) {
          idx_cur = idx_nxt;
          chunk_base += 32;
          idx_nxt = (chunk_base + 32 + lane < ee) ? p.indices[chunk_base + 32 + lane] : 0;
        }
        if (p.bulk) {
          const unsigned q = (unsigned)(e - eb);
          const int slot = (int)(q % kDepth);
          mbar_wait(&wbar[slot], (q / kDepth) & 1u);
          acc.add_smem(wring + slot * SM::kSeg, lane);
          __syncwarp();                                             // every lane has read the slot
          const int64_t pe = e + kDepth;                            // stream position that reuses it
          if (pe < ee) {
            const int poff = (int)(pe - chunk_base);                // < 32 + kDepth
            if (lane == (poff & 31)) {
              const int m = poff < 32 ? idx_cur : idx_nxt;
              fence_proxy_async();                                  // generic-proxy reads before the async write
              mbar_expect_tx(&wbar[slot], seg_bytes);
              bulk_copy(wring + slot * SM::kSeg, xbase + (int64_t)m * pitch, seg_bytes, &wbar[slot], pol);
            }
          }
        } else {
          const int m = __shfl_sync(0xffffffffu, idx_cur, (int)(e - chunk_base));
          acc.add_global(xbase + (int64_t)m * pitch, lane, valid);
        }
      }

      // ------------------------------------------------------------------ row `row` is complete
      const int il = (int)(row - i0);
      if (MODE == MODE_FIRST) {
        if (kU16) {
          const double bo = row_bound(p.out_bound, row);
          const double inv = bo > 0.0 ? 65535.0 / bo : 0.0;
#pragma unroll
          for (int j = 0; j < Acc<E, TC>::kCols; ++j) {
            const int cl = Acc<E, TC>::col(j, lane);
            const int64_t c = c0 + cl;
            double q = c < p.L ? rint(acc.val(j) * row_bound(p.in_unit, c) * inv) : 0.0;
            if (!(q > 0.0)) q = 0.0;
            if (q > 65535.0) q = 65535.0;
            tile[cl * SM::kPitch + il] = (TileT)(unsigned)q;
          }
        } else {
          const double gi = p.g[row];
#pragma unroll
          for (int j = 0; j < Acc<E, TC>::kCols; ++j)
            tile[Acc<E, TC>::col(j, lane) * SM::kPitch + il] = (TileT)(acc.val(j) * gi);
        }
      } else if (MODE == MODE_FINAL) {
        if (kU16) {
#pragma unroll
          for (int j = 0; j < Acc<E, TC>::kCols; ++j) tile[Acc<E, TC>::col(j, lane) * SM::kPitch + il] = (TileT)acc.val(j);
        } else {
          const double gi = p.g[row];
#pragma unroll
          for (int j = 0; j < Acc<E, TC>::kCols; ++j)
            tile[Acc<E, TC>::col(j, lane) * SM::kPitch + il] = (TileT)(acc.val(j) * gi);
        }
      } else {
        // symmetric FINAL: element (row, c) for c >= row, also stored at (c, row)
        const double gi = p.g[row] * p.epi.coef;
        double* orow = reinterpret_cast<double*>(p.OUT) + row * p.ldo;
#pragma unroll
        for (int j = 0; j < Acc<E, TC>::kCols; ++j) {
          const int64_t c = c0 + Acc<E, TC>::col(j, lane);
          if (c >= p.L || c < row) continue;
          uint32_t cnt = 0u;
          if (p.counts) cnt = load_count(p.counts, row * p.ld_counts + c, p.counts32);
          double v;
          if (kU16)
            v = gi * p.g_col[c] * (acc.val(j) * row_bound(p.in_unit, c) + (p.add_counts ? (double)cnt : 0.0));
          else
            v = acc.val(j) * gi;
          if (p.use_evidence) v *= evidence_factor(cnt);
          else if (p.epi.evidence) v *= evidence_factor(__ldcs(p.epi.evidence + row * p.epi.ld_evidence + c));
          if (c == row) v = 1.0; else if (v > omax) omax = v;
          if (p.epi.s_old) {
            const double d = fabs(v - __ldcs(p.epi.s_old + row * p.epi.ld_s_old + c));
            if (d > dmax) dmax = d;
          }
          __stcs(orow + c, v);
          if (c > row) __stcs(reinterpret_cast<double*>(p.OUT) + c * p.ldo + row, v);
        }
      }
      acc.clear();
    }
  }

  if (MODE != MODE_FINAL_SYM) {
    __syncthreads();
    const int64_t i = i0 + lane;
    for (int cl = warp; cl < TC; cl += kWarps) {
      const int64_t r = c0 + cl;
      if (r >= p.L || i >= p.row_end) continue;
      if (MODE == MODE_FIRST) {
        if (kU16) reinterpret_cast<uint16_t*>(p.OUT)[r * p.ldo + i] = (uint16_t)tile[cl * SM::kPitch + lane];
        else __stcs(reinterpret_cast<double*>(p.OUT) + r * p.ldo + i, (double)tile[cl * SM::kPitch + lane]);
        continue;
      }
      uint32_t cnt = 0u;
      if (p.counts) cnt = load_count(p.counts, r * p.ld_counts + i, p.counts32);
      double v;
      if (kU16)
        v = p.g[i] * p.g_col[r] * ((double)tile[cl * SM::kPitch + lane] * row_bound(p.in_unit, r) +
                                  (p.add_counts ? (double)cnt : 0.0));
      else
        v = (double)tile[cl * SM::kPitch + lane];
      v *= p.epi.coef;
      // the epilogue streams (evidence, prior, S_old, the result) are touched once: evict-first
      // loads/stores keep them from pushing the gathered panel of X out of L2
      if (p.use_evidence) v *= evidence_factor(cnt);
      else if (p.epi.evidence) v *= evidence_factor(__ldcs(p.epi.evidence + r * p.epi.ld_evidence + i));
      if (p.epi.prior) v = (1.0 - p.epi.lambda) * v + p.epi.lambda * __ldcs(p.epi.prior + r * p.epi.ld_prior + i);
      if (r + p.epi.diag_offset == i) v = 1.0; else if (v > omax) omax = v;
      if (p.epi.s_old) {
        const double d = fabs(v - __ldcs(p.epi.s_old + r * p.epi.ld_s_old + i));
        if (d > dmax) dmax = d;                  // NaN compares false: ignored like SimRank.py:74
      }
      __stcs(reinterpret_cast<double*>(p.OUT) + r * p.ldo + i, v);
    }
  }
  if (MODE != MODE_FIRST) {
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
// Fixed-point source operand of the u16 gather.  unit[r] = max_k V[r, k] / 65535 over the row (the
// element (r, r + zero_diag_offset) excluded: S = I + S_off), then XT[k, r] = rint(V[r, k] / unit[r]):
// the TRANSPOSED matrix with one scale per column -- for a symmetric S this is S_off itself scaled by
// its column maxima, which is what lets the gather sum whole columns as integers.
constexpr int QT = 64;
__global__ void __launch_bounds__(256)
row_unit_kernel(const double* __restrict__ V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                double* __restrict__ unit) {
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
    if (lane == 0) unit[r] = m / 65535.0;
  }
}
__global__ void __launch_bounds__(256)
quantize_transpose_u16_kernel(const double* __restrict__ V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                              const double* __restrict__ unit, uint16_t* __restrict__ XT, int64_t ldxt) {
  __shared__ uint16_t t[QT][QT + 2];
  const int64_t r0 = (int64_t)blockIdx.y * QT, k0 = (int64_t)blockIdx.x * QT;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
  for (int rr = ty; rr < QT; rr += 4) {
    const int64_t r = r0 + rr, k = k0 + tx;
    unsigned q = 0u;
    if (r < R && k < K) {
      const double u = unit[r];
      const double v = (zero_diag_offset >= 0 && k == r + zero_diag_offset) ? 0.0 : V[r * ldv + k];
      double x = u > 0.0 ? rint(v / u) : 0.0;
      if (!(x > 0.0)) x = 0.0;
      if (x > 65535.0) x = 65535.0;
      q = (unsigned)x;
    }
    t[tx][rr] = (uint16_t)q;
  }
  __syncthreads();
  for (int kk = ty; kk < QT; kk += 4) {
    const int64_t k = k0 + kk, r = r0 + tx;
    if (k < K && r < ldxt) XT[k * ldxt + r] = r < R ? t[kk][tx] : (uint16_t)0;
  }
}

template <typename E, int TC, int MODE>
static int launch_one(const Params& p, dim3 grid, cudaStream_t st) {
  typedef Smem<E, TC, MODE> SM;
  auto kern = csr_gather_kernel<E, TC, MODE>;
  SRK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
  kern<<<grid, kThreads, SM::kBytes, st>>>(p);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

}  // namespace gat
}  // namespace srk

using namespace srk;

extern "C" int srk_csr_half(const srk_csr_args* a, void* stream) {
  SRK_REQUIRE(a, "null args");
  SRK_REQUIRE(a->indptr && a->indices && a->g && a->X && a->OUT, "null pointer");
  SRK_REQUIRE(0 <= a->row_begin && a->row_begin <= a->row_end && a->row_end <= a->M, "row range");
  // OUT is addressed as OUT[c * ldo + i] for i in [row_begin, row_end) only: a caller that stores just
  // those columns passes the address of (virtual) column 0, i.e. its buffer minus row_begin elements
  SRK_REQUIRE(a->L >= 0 && a->ldx >= a->L && a->ldo >= a->row_end - a->row_begin, "leading dimensions");
  SRK_REQUIRE(a->elem == SRK_ELEM_F64 || a->elem == SRK_ELEM_U16, "elem must be SRK_ELEM_F64 or SRK_ELEM_U16");
  SRK_REQUIRE(a->mode == SRK_CSR_FIRST || a->mode == SRK_CSR_FINAL, "mode must be SRK_CSR_FIRST or SRK_CSR_FINAL");
  SRK_REQUIRE(a->counts_bits == 0 || a->counts_bits == 16 || a->counts_bits == 32, "counts_bits must be 16 or 32");
  SRK_REQUIRE(!(a->add_counts || a->use_evidence) || a->counts, "counts missing");
  SRK_REQUIRE(!(a->use_evidence && a->epi.evidence), "evidence given twice (counts and epi.evidence)");
  const bool sym = a->mode == SRK_CSR_FINAL && a->symmetric;
  if (sym)
    SRK_REQUIRE(a->row_begin == 0 && a->row_end == a->M && a->L == a->M && a->epi.prior == nullptr &&
                    a->epi.diag_offset == 0 && a->ldo >= a->L,
                "the symmetric second half needs the whole square problem and no prior");
  if (a->elem == SRK_ELEM_U16 && a->mode == SRK_CSR_FINAL) SRK_REQUIRE(a->g_col, "u16 FINAL needs g_col");
  if (a->row_end == a->row_begin || a->L == 0) return SRK_OK;

  gat::Params p;
  memset(&p, 0, sizeof(p));
  p.indptr = a->indptr; p.indices = a->indices; p.g = a->g;
  p.row_begin = a->row_begin; p.row_end = a->row_end;
  p.X = a->X; p.ldx = a->ldx; p.L = a->L; p.OUT = a->OUT; p.ldo = a->ldo;
  p.in_unit = a->in_unit; p.out_bound = a->out_bound; p.g_col = a->g_col;
  p.counts = a->counts; p.ld_counts = a->ld_counts; p.counts32 = a->counts_bits == 32;
  p.add_counts = a->add_counts; p.use_evidence = a->use_evidence;
  if (a->mode == SRK_CSR_FINAL) { p.epi = to_dev(a->epi); p.maxdiff = a->epi.maxdiff; p.maxoff = a->epi.maxoff; }
  const int64_t esz = a->elem == SRK_ELEM_U16 ? 2 : 8;
  p.bulk = ((uintptr_t)a->X % 16 == 0) && ((a->ldx * esz) % 16 == 0);
  { const char* e = getenv("SRK_CSR_FLAGS"); p.flags = e ? atoi(e) : 0; if (p.flags & 2) p.bulk = 0; }

  // Panel width = columns of X per CTA = what all CTAs of a grid column gather from; the panel (rows
  // of X x segment bytes, held once per L2 die) has to survive in L2 next to the streams of the
  // epilogue.  float64: 128 columns (1 KB segments) for the first half, 64 for the second, which also
  // streams S_old (with 128 its panel fell out of L2: 93 ms against 42 ms at n = 32768).  uint16: 256
  // columns (512 B segments).
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = a->row_end - a->row_begin;
  const int mode = a->mode == SRK_CSR_FIRST ? gat::MODE_FIRST : (sym ? gat::MODE_FINAL_SYM : gat::MODE_FINAL);
  const int tc = a->elem == SRK_ELEM_U16 ? 256 : (mode == gat::MODE_FIRST ? 128 : 64);
  const int64_t gx = (rows + gat::TI - 1) / gat::TI, gy = (a->L + tc - 1) / tc;
  SRK_REQUIRE(gy <= 65535, "too many column panels");
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (a->elem == SRK_ELEM_U16) {
    switch (mode) {
      case gat::MODE_FIRST: return gat::launch_one<uint16_t, 256, gat::MODE_FIRST>(p, grid, st);
      case gat::MODE_FINAL: return gat::launch_one<uint16_t, 256, gat::MODE_FINAL>(p, grid, st);
      default: return gat::launch_one<uint16_t, 256, gat::MODE_FINAL_SYM>(p, grid, st);
    }
  }
  switch (mode) {
    case gat::MODE_FIRST: return gat::launch_one<double, 128, gat::MODE_FIRST>(p, grid, st);
    case gat::MODE_FINAL: return gat::launch_one<double, 64, gat::MODE_FINAL>(p, grid, st);
    default: return gat::launch_one<double, 64, gat::MODE_FINAL_SYM>(p, grid, st);
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
                                     uint16_t* XT, int64_t ldxt, double* unit, void* stream) {
  SRK_REQUIRE(V && XT && unit, "null pointer");
  SRK_REQUIRE(ldv >= K && ldxt >= R && ldxt % 8 == 0 && ((uintptr_t)XT % 16) == 0,
              "XT must be 16-byte aligned with ldxt a multiple of 8 and >= R");
  if (R == 0 || K == 0) return SRK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  gat::row_unit_kernel<<<(unsigned)R, 256, 0, st>>>(V, ldv, R, K, zero_diag_offset, unit);
  dim3 grid((unsigned)((K + gat::QT - 1) / gat::QT), (unsigned)((ldxt + gat::QT - 1) / gat::QT));
  SRK_REQUIRE(grid.y <= 65535, "too many row tiles");
  gat::quantize_transpose_u16_kernel<<<grid, 256, 0, st>>>(V, ldv, R, K, zero_diag_offset, unit, XT, ldxt);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}
