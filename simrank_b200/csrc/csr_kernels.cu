// The small preprocessing / retrieval kernels around the half-products (evidence counts, row
// spread, CSR -> dense pattern, fixed-point slicing, top-k).  The CSR half-products themselves are
// in csr_gather.cu, the tensor-core ones in dense_i8x2.cu.
#include <stdlib.h>

#include "common.cuh"

namespace srk {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

// ------------------------------------------------------------------------------------------
// cnt[i, j] = |N(i) & N(j)| by merging the two sorted neighbour lists; one thread per pair.
__global__ void evidence_counts_kernel(const int64_t* __restrict__ indptr,
                                       const int32_t* __restrict__ indices,
                                       const uint8_t* __restrict__ dead, int64_t M,
                                       int64_t row_begin, int64_t row_end,
                                       uint8_t* __restrict__ counts, int64_t ldc) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = row_begin + (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= row_end || j >= M) return;
  unsigned c = 0;
  if (!(dead && (dead[i] || dead[j]))) {
    int64_t a = indptr[i], ae = indptr[i + 1], b = indptr[j], be = indptr[j + 1];
    while (a < ae && b < be) {
      const int32_t x = indices[a], y = indices[b];
      c += (x == y);
      a += (x <= y);
      b += (y <= x);
    }
  }
  counts[(i - row_begin) * ldc + j] = (uint8_t)min(c, 255u);
}

// ------------------------------------------------------------------------------------------
// Two-pass sample variance of the nonzero entries of each CSR row (pandas nanvar, ddof=1).
__global__ void row_spread_kernel(const int64_t* __restrict__ indptr, const double* __restrict__ vals,
                                  const double* __restrict__ g, int64_t M, double* __restrict__ spread) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t beg = indptr[i], end = indptr[i + 1];
  double sum = 0.0;
  int64_t cnt = 0;
  for (int64_t e = beg; e < end; ++e) {
    const double x = vals ? vals[e] : g[i];
    if (x != 0.0 && x == x) { sum += x; ++cnt; }
  }
  double var = 0.0;
  if (cnt >= 2) {
    const double mean = sum / (double)cnt;
    double sq = 0.0;
    for (int64_t e = beg; e < end; ++e) {
      const double x = vals ? vals[e] : g[i];
      if (x != 0.0 && x == x) { const double d = mean - x; sq += d * d; }
    }
    var = sq / (double)(cnt - 1);
    if (var != var) var = 0.0;
  }
  spread[i] = exp(-var);
}

// ------------------------------------------------------------------------------------------
// Edge list -> CSR on the device (the pivot + row scatter of SimRank.py:50-52, 199-200):
//   1. degree histogram (one atomicAdd per edge)                          edge_degree_kernel
//   2. exclusive scan of the degrees -> indptr (three small kernels)      scan_*_kernel
//   3. scatter of the column indices into their row, in arrival order     edge_scatter_kernel
//   4. per row: a K-bit bitmap in shared memory is set from the row's columns (a bit that is already
//      set is a duplicate (row, column) pair: the reference's pivot raises on those), its popcount
//      prefix gives every set bit its rank, and the columns are written back in increasing order --
//      a counting sort that needs no comparison and does not care how long the row is.
constexpr int SCAN_THREADS = 1024, SCAN_ITEMS = 4;
__global__ void edge_degree_kernel(const int32_t* __restrict__ rows, int64_t m, int64_t M, int32_t* __restrict__ deg,
                                   int32_t* __restrict__ bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int32_t r = rows[e];
  if (r < 0 || r >= M) { atomicOr(bad, 2); return; }
  atomicAdd(deg + r, 1);
}
__global__ void __launch_bounds__(SCAN_THREADS)
scan_blocks_kernel(const int32_t* __restrict__ deg, int64_t M, int64_t* __restrict__ indptr, int64_t* __restrict__ block_sum) {
  __shared__ int64_t warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = ((int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  int64_t v[SCAN_ITEMS], run = 0;
#pragma unroll
  for (int x = 0; x < SCAN_ITEMS; ++x) { v[x] = run; run += base + x < M ? (int64_t)deg[base + x] : 0; }
  int64_t inc = run;                                         // inclusive scan of the thread totals over the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int64_t w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    warp_tot[lane] = w;                                       // inclusive over the warps
  }
  __syncthreads();
  const int64_t before = (warp ? warp_tot[warp - 1] : 0) + inc - run;
#pragma unroll
  for (int x = 0; x < SCAN_ITEMS; ++x)
    if (base + x < M) indptr[base + x] = before + v[x];      // block-local exclusive prefix
  if (threadIdx.x == SCAN_THREADS - 1) block_sum[blockIdx.x] = warp_tot[SCAN_THREADS / 32 - 1];
}
__global__ void scan_sums_kernel(int64_t* __restrict__ block_sum, int64_t blocks) {   // one thread: a few thousand adds
  if (blockIdx.x || threadIdx.x) return;
  int64_t run = 0;
  for (int64_t b = 0; b < blocks; ++b) { const int64_t t = block_sum[b]; block_sum[b] = run; run += t; }
}
__global__ void scan_add_kernel(int64_t* __restrict__ indptr, int64_t M, int64_t m, const int64_t* __restrict__ block_sum) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) indptr[i] += block_sum[i / (SCAN_THREADS * SCAN_ITEMS)];
  if (i == M) indptr[M] = m;
}
__global__ void edge_scatter_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int64_t m, int64_t M,
                                    int64_t K, const int64_t* __restrict__ indptr, int32_t* __restrict__ cursor,
                                    int32_t* __restrict__ tmp, int32_t* __restrict__ bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int32_t r = rows[e], c = cols[e];
  if (r < 0 || r >= M) return;
  if (c < 0 || c >= K) { atomicOr(bad, 2); return; }
  tmp[indptr[r] + atomicAdd(cursor + r, 1)] = c;
}
constexpr int SORT_THREADS = 256;
__global__ void __launch_bounds__(SORT_THREADS)
row_bitmap_sort_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ tmp, int64_t M, int64_t K,
                       int32_t* __restrict__ indices, int32_t* __restrict__ bad) {
  extern __shared__ uint32_t bm[];                            // [words] bitmap, then [words] popcount prefix
  __shared__ uint32_t warp_tot[SORT_THREADS / 32];
  __shared__ uint32_t carry;
  const int words = (int)((K + 31) / 32);
  uint32_t* pre = bm + words;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r = blockIdx.x; r < M; r += gridDim.x) {
    const int64_t beg = indptr[r], end = indptr[r + 1];
    if (end - beg <= 1) {                                     // nothing to sort
      if (end > beg && threadIdx.x == 0) indices[beg] = tmp[beg];
      continue;
    }
    for (int w = threadIdx.x; w < words; w += SORT_THREADS) bm[w] = 0u;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (int64_t e = beg + threadIdx.x; e < end; e += SORT_THREADS) {
      const int32_t c = tmp[e];
      const uint32_t bit = 1u << (c & 31);
      if (atomicOr(&bm[c >> 5], bit) & bit) atomicOr(bad, 1);  // duplicate (row, column) pair
    }
    __syncthreads();
    // exclusive prefix of the word popcounts, SORT_THREADS words per round
    for (int w0 = 0; w0 < words; w0 += SORT_THREADS) {
      const int w = w0 + threadIdx.x;
      const uint32_t cnt = w < words ? __popc(bm[w]) : 0u;
      uint32_t inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      if (lane == 31) warp_tot[warp] = inc;
      __syncthreads();
      uint32_t before = carry;
      for (int x = 0; x < warp; ++x) before += warp_tot[x];
      if (w < words) pre[w] = before + inc - cnt;
      __syncthreads();
      if (threadIdx.x == SORT_THREADS - 1) carry = before + inc;
      __syncthreads();
    }
    for (int w = threadIdx.x; w < words; w += SORT_THREADS) {
      uint32_t b = bm[w];
      int64_t o = beg + pre[w];
      while (b) {
        const int t = __ffs(b) - 1;
        indices[o++] = w * 32 + t;
        b &= b - 1;
      }
    }
    __syncthreads();                                          // bm / pre are reused by the next row
  }
}

// ------------------------------------------------------------------------------------------
__global__ void csr_scatter_u8_kernel(const int64_t* __restrict__ indptr,
                                      const int32_t* __restrict__ indices, int64_t row_begin,
                                      int64_t row_end, int64_t K, uint8_t* __restrict__ A8, int64_t lda) {
  const int64_t i = row_begin + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (i >= row_end) return;
  for (int64_t e = indptr[i] + lane; e < indptr[i + 1]; e += 32) {
    const int64_t c = indices[e];
    if (c < K) A8[(i - row_begin) * lda + c] = 1;
  }
}

// ------------------------------------------------------------------------------------------
// f64 -> NS uint8 planes; one thread per 16 consecutive columns (16 B store per plane).
template <int NS>
__global__ void slice_rows_kernel(const double* __restrict__ V, int64_t ldv, int64_t R, int64_t K,
                                  const srk_rowbound rowbound, int64_t zero_diag_offset,
                                  uint8_t* __restrict__ planes, int64_t ldp, int64_t plane_stride) {
  const int64_t chunks = (ldp + 15) / 16;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * chunks) return;
  const int64_t r = t / chunks, k0 = (t % chunks) * 16;
  const double qmax = (double)((1ull << (8 * NS)) - 1ull);
  const double b = row_bound(rowbound, r);
  const double scale = (b > 0.0) ? (qmax + 1.0) / b : 0.0;
  uint32_t w[NS][4];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int x = 0; x < 4; ++x) w[s][x] = 0u;
#pragma unroll
  for (int x = 0; x < 16; ++x) {
    const int64_t k = k0 + x;
    double v = (k < K) ? V[r * ldv + k] : 0.0;
    if (zero_diag_offset >= 0 && k == r + zero_diag_offset) v = 0.0;
    double q = rint(v * scale);
    if (!(q > 0.0)) q = 0.0;                 // negatives and NaN -> 0
    if (q > qmax) q = qmax;
    const unsigned long long qi = (unsigned long long)q;
#pragma unroll
    for (int s = 0; s < NS; ++s)
      w[s][x >> 2] |= (uint32_t)((qi >> (8 * (NS - 1 - s))) & 0xffull) << (8 * (x & 3));
  }
#pragma unroll
  for (int s = 0; s < NS; ++s)
    *reinterpret_cast<uint4*>(planes + s * plane_stride + r * ldp + k0) =
        make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
}

// ------------------------------------------------------------------------------------------
// Planes with the exact per-row bound (srk_slice_rows_max_f64): one CTA per row, two passes over
// the row (the second one is served by L2: a row is at most a few hundred KB).
constexpr int SLICE_THREADS = 512;
__device__ __forceinline__ void load16(const double* p, int64_t k0, int64_t K, bool vec, double (&v)[16]) {
  if (vec && k0 + 16 <= K) {
#pragma unroll
    for (int x = 0; x < 16; x += 2) {
      const double2 d = *reinterpret_cast<const double2*>(p + k0 + x);
      v[x] = d.x; v[x + 1] = d.y;
    }
  } else {
#pragma unroll
    for (int x = 0; x < 16; ++x) v[x] = (k0 + x < K) ? p[k0 + x] : 0.0;
  }
}
template <int NS>
__global__ void __launch_bounds__(SLICE_THREADS)
slice_rows_max_kernel(const double* __restrict__ V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                      uint8_t* __restrict__ planes, int64_t ldp, int64_t plane_stride,
                      double* __restrict__ bound_out) {
  __shared__ double red[SLICE_THREADS / 32];
  __shared__ double row_max;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double qmax = (double)((1ull << (8 * NS)) - 1ull);
  const int64_t chunks = (ldp + 15) / 16;
  for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
    const double* row = V + r * ldv;
    const bool vec = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
    const int64_t kd = zero_diag_offset >= 0 ? r + zero_diag_offset : -1;
    double m = 0.0;
    for (int64_t c = threadIdx.x; c * 16 < K; c += SLICE_THREADS) {
      double v[16];
      load16(row, c * 16, K, vec, v);
#pragma unroll
      for (int x = 0; x < 16; ++x)
        if (c * 16 + x != kd && v[x] > m) m = v[x];          // NaN and negatives never win
    }
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (warp == 0) {
      m = lane < SLICE_THREADS / 32 ? red[lane] : 0.0;
      m = warp_max(m);
      if (lane == 0) {
        row_max = m;
        if (bound_out) bound_out[r] = m > 0.0 ? m * ((qmax + 1.0) / qmax) : 0.0;
      }
    }
    __syncthreads();
    m = row_max;
    const double scale = m > 0.0 ? qmax / m : 0.0;
    for (int64_t c = threadIdx.x; c < chunks; c += SLICE_THREADS) {
      const int64_t k0 = c * 16;
      double v[16];
      load16(row, k0, K, vec, v);
      uint32_t w[NS][4];
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int x = 0; x < 4; ++x) w[s][x] = 0u;
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        double q = (k0 + x == kd) ? 0.0 : rint(v[x] * scale);
        if (!(q > 0.0)) q = 0.0;
        if (q > qmax) q = qmax;
        const unsigned long long qi = (unsigned long long)q;
#pragma unroll
        for (int s = 0; s < NS; ++s)
          w[s][x >> 2] |= (uint32_t)((qi >> (8 * (NS - 1 - s))) & 0xffull) << (8 * (x & 3));
      }
#pragma unroll
      for (int s = 0; s < NS; ++s)
        *reinterpret_cast<uint4*>(planes + s * plane_stride + r * ldp + k0) =
            make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
    __syncthreads();                                          // row_max / red are reused by the next row
  }
}

// One-pass variant: the row maxima come as keys from the FINAL epilogue (srk_x2_args.rowmax_hi).
template <int NS>
__global__ void slice_rows_key_kernel(const double* __restrict__ V, int64_t ldv, int64_t R, int64_t K,
                                      int64_t zero_diag_offset, const uint32_t* __restrict__ key,
                                      uint8_t* __restrict__ planes, int64_t ldp, int64_t plane_stride,
                                      double* __restrict__ bound_out) {
  const int64_t chunks = (ldp + 15) / 16;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * chunks) return;
  const int64_t r = t / chunks, k0 = (t % chunks) * 16;
  const double qmax = (double)((1ull << (8 * NS)) - 1ull);
  const uint32_t kr = key[r];
  const double m = kr > 1u ? __longlong_as_double((long long)kr << 32) : 0.0;
  const double scale = m > 0.0 ? qmax / m : 0.0;
  if (k0 == 0 && bound_out) bound_out[r] = m > 0.0 ? m * ((qmax + 1.0) / qmax) : 0.0;
  const double* row = V + r * ldv;
  double v[16];
  load16(row, k0, K, (reinterpret_cast<uintptr_t>(row) & 15) == 0, v);
  const int64_t kd = zero_diag_offset >= 0 ? r + zero_diag_offset : -1;
  uint32_t w[NS][4];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int x = 0; x < 4; ++x) w[s][x] = 0u;
#pragma unroll
  for (int x = 0; x < 16; ++x) {
    double q = (k0 + x == kd) ? 0.0 : rint(v[x] * scale);
    if (!(q > 0.0)) q = 0.0;
    if (q > qmax) q = qmax;
    const unsigned long long qi = (unsigned long long)q;
#pragma unroll
    for (int s = 0; s < NS; ++s)
      w[s][x >> 2] |= (uint32_t)((qi >> (8 * (NS - 1 - s))) & 0xffull) << (8 * (x & 3));
  }
#pragma unroll
  for (int s = 0; s < NS; ++s)
    *reinterpret_cast<uint4*>(planes + s * plane_stride + r * ldp + k0) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
}

// ------------------------------------------------------------------------------------------
// Row-wise top-k by k rounds of arg-max in the total order (value desc, column asc); NaN last.
constexpr int TOPK_THREADS = 256;
__device__ __forceinline__ bool topk_before(double va, int ia, double vb, int ib) {
  // true when (va, ia) ranks strictly before (vb, ib)
  const bool na = va != va, nb = vb != vb;
  if (na != nb) return nb;
  if (!na && va != vb) return va > vb;
  return ia < ib;
}
__global__ void __launch_bounds__(TOPK_THREADS)
topk_rows_kernel(const double* __restrict__ S, int64_t lds, int64_t n, int k,
                 int32_t* __restrict__ idx, double* __restrict__ vals) {
  __shared__ double sv[TOPK_THREADS / 32];
  __shared__ int si[TOPK_THREADS / 32];
  __shared__ double last_v;
  __shared__ int last_i;
  const double* row = S + (int64_t)blockIdx.x * lds;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { last_i = -1; last_v = 0.0; }
  __syncthreads();
  for (int round = 0; round < k; ++round) {
    const int li = last_i;
    const double lv = last_v;
    double bv = 0.0;
    int bi = -1;
    for (int64_t c = threadIdx.x; c < n; c += TOPK_THREADS) {
      const double v = row[c];
      if (li >= 0 && !topk_before(lv, li, v, (int)c)) continue;   // already emitted
      if (bi < 0 || topk_before(v, (int)c, bv, bi)) { bv = v; bi = (int)c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || topk_before(ov, oi, bv, bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = (lane < TOPK_THREADS / 32) ? sv[lane] : 0.0;
      bi = (lane < TOPK_THREADS / 32) ? si[lane] : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || topk_before(ov, oi, bv, bi))) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        idx[(int64_t)blockIdx.x * k + round] = bi;
        vals[(int64_t)blockIdx.x * k + round] = bv;
        last_i = bi; last_v = bv;
      }
    }
    __syncthreads();
  }
}

__global__ void set_identity_kernel(double* __restrict__ S, int64_t lds, int64_t R, int64_t n,
                                    int64_t diag_offset) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * n) return;
  const int64_t r = t / n, c = t % n;
  S[r * lds + c] = (c == r + diag_offset) ? 1.0 : 0.0;
}

}  // namespace srk

// =============================================================================== C ABI
using namespace srk;

extern "C" int srk_abi_version(void) { return SRK_ABI_VERSION; }
extern "C" const char* srk_last_error(void) { return error_buffer(); }

extern "C" int srk_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  SRK_CUDA_OK(cudaGetDevice(&dev));
  SRK_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SRK_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

extern "C" int srk_csr_evidence_counts(const int64_t* indptr, const int32_t* indices,
                                       const uint8_t* dead, int64_t M, int64_t row_begin,
                                       int64_t row_end, uint8_t* counts, int64_t ldc, void* stream) {
  SRK_REQUIRE(indptr && indices && counts, "null pointer");
  SRK_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= M && ldc >= M, "shape");
  if (row_end == row_begin) return SRK_OK;
  dim3 block(32, 8);
  dim3 grid((unsigned)((M + 31) / 32), (unsigned)((row_end - row_begin + 7) / 8));
  SRK_REQUIRE(grid.y <= 65535, "row shard too tall for one launch");
  evidence_counts_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(indptr, indices, dead, M, row_begin,
                                                                  row_end, counts, ldc);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_csr_row_spread(const int64_t* indptr, const double* vals, const double* g,
                                  int64_t M, double* spread, void* stream) {
  SRK_REQUIRE(indptr && spread && (vals || g), "null pointer");
  if (M == 0) return SRK_OK;
  row_spread_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(indptr, vals, g, M, spread);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_csr_to_dense_u8(const int64_t* indptr, const int32_t* indices, int64_t row_begin,
                                   int64_t row_end, int64_t K, uint8_t* A8, int64_t lda, void* stream) {
  SRK_REQUIRE(indptr && indices && A8, "null pointer");
  SRK_REQUIRE(row_begin <= row_end && lda >= K, "shape");
  if (row_end == row_begin) return SRK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SRK_CUDA_OK(cudaMemsetAsync(A8, 0, (size_t)(row_end - row_begin) * lda, st));
  const int64_t threads = (row_end - row_begin) * 32;
  csr_scatter_u8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(indptr, indices, row_begin,
                                                                          row_end, K, A8, lda);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" size_t srk_edges_to_csr_workspace(int64_t m, int64_t M) {
  const size_t blocks = (size_t)((M + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS)) + 1;
  // degree + cursor (int32 [M] each), arrival-order columns (int32 [m]), block sums (int64), each 256-byte aligned
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  return up((size_t)M * 4) * 2 + up((size_t)m * 4) + up(blocks * 8) + 256;
}

extern "C" int srk_edges_to_csr(const int32_t* rows, const int32_t* cols, int64_t m, int64_t M, int64_t K,
                                int64_t* indptr, int32_t* indices, int32_t* status, void* workspace,
                                size_t workspace_bytes, void* stream) {
  SRK_REQUIRE(indptr && status && (m == 0 || (rows && cols && indices)), "null pointer");
  SRK_REQUIRE(m >= 0 && M >= 0 && K >= 0 && m < (1ll << 31) && M < (1ll << 31), "shape");
  SRK_REQUIRE(workspace_bytes >= srk_edges_to_csr_workspace(m, M) && (M == 0 || workspace), "workspace too small");
  const size_t bitmap_bytes = (size_t)((K + 31) / 32) * 8;
  SRK_REQUIRE(bitmap_bytes <= 200 * 1024, "too many columns for the shared-memory row bitmap (K <= 819200)");
  cudaStream_t st = (cudaStream_t)stream;
  SRK_CUDA_OK(cudaMemsetAsync(status, 0, 4, st));
  if (M == 0) { SRK_CUDA_OK(cudaMemsetAsync(indptr, 0, 8, st)); return SRK_OK; }
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  int32_t* deg = reinterpret_cast<int32_t*>(ws);
  int32_t* cursor = reinterpret_cast<int32_t*>(ws + up((size_t)M * 4));
  int32_t* tmp = reinterpret_cast<int32_t*>(ws + 2 * up((size_t)M * 4));
  int64_t* block_sum = reinterpret_cast<int64_t*>(ws + 2 * up((size_t)M * 4) + up((size_t)m * 4));
  SRK_CUDA_OK(cudaMemsetAsync(deg, 0, 2 * up((size_t)M * 4), st));            // degrees and cursors
  const unsigned eb = (unsigned)((m + 255) / 256);
  if (m) edge_degree_kernel<<<eb, 256, 0, st>>>(rows, m, M, deg, status);
  const int64_t blocks = (M + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS);
  scan_blocks_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(deg, M, indptr, block_sum);
  scan_sums_kernel<<<1, 32, 0, st>>>(block_sum, blocks);
  scan_add_kernel<<<(unsigned)((M + 1 + 255) / 256), 256, 0, st>>>(indptr, M, m, block_sum);
  if (m) {
    edge_scatter_kernel<<<eb, 256, 0, st>>>(rows, cols, m, M, K, indptr, cursor, tmp, status);
    int dev = 0, sms = 0;
    SRK_CUDA_OK(cudaGetDevice(&dev));
    SRK_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SRK_CUDA_OK(cudaFuncSetAttribute(row_bitmap_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bitmap_bytes));
    const int64_t per_sm = bitmap_bytes ? (int64_t)(200 * 1024 / (bitmap_bytes + 1024)) : 8;
    int64_t grid = (int64_t)sms * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    if (grid > M) grid = M;
    row_bitmap_sort_kernel<<<(unsigned)grid, SORT_THREADS, bitmap_bytes, st>>>(indptr, tmp, M, K, indices, status);
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_slice_rows_f64(const double* V, int64_t ldv, int64_t R, int64_t K,
                                  const srk_rowbound* rowbound, int64_t zero_diag_offset, int ns,
                                  uint8_t* planes, int64_t ldp, int64_t plane_stride, void* stream) {
  SRK_REQUIRE(V && rowbound && planes, "null pointer");
  SRK_REQUIRE(ldp % 16 == 0 && ldp >= K && ldv >= K, "ldp must be a multiple of 16 and >= K");
  SRK_REQUIRE(((uintptr_t)planes % 16) == 0 && plane_stride % 16 == 0, "planes must be 16-byte aligned");
  if (R == 0 || K == 0) return SRK_OK;
  const int64_t threads = R * ((ldp + 15) / 16);
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (ns) {
    case 1: slice_rows_kernel<1><<<blocks, 256, 0, st>>>(V, ldv, R, K, *rowbound, zero_diag_offset, planes, ldp, plane_stride); break;
    case 2: slice_rows_kernel<2><<<blocks, 256, 0, st>>>(V, ldv, R, K, *rowbound, zero_diag_offset, planes, ldp, plane_stride); break;
    case 3: slice_rows_kernel<3><<<blocks, 256, 0, st>>>(V, ldv, R, K, *rowbound, zero_diag_offset, planes, ldp, plane_stride); break;
    case 4: slice_rows_kernel<4><<<blocks, 256, 0, st>>>(V, ldv, R, K, *rowbound, zero_diag_offset, planes, ldp, plane_stride); break;
    default: return srk::fail(SRK_ERR_INVALID, "invalid argument: %s", "ns must be 1..4");
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_slice_rows_max_f64(const double* V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                                      int ns, uint8_t* planes, int64_t ldp, int64_t plane_stride,
                                      double* bound_out, void* stream) {
  SRK_REQUIRE(V && planes, "null pointer");
  SRK_REQUIRE(ldp % 16 == 0 && ldp >= K && ldv >= K, "ldp must be a multiple of 16 and >= K");
  SRK_REQUIRE(((uintptr_t)planes % 16) == 0 && plane_stride % 16 == 0, "planes must be 16-byte aligned");
  if (R == 0 || K == 0) return SRK_OK;
  int dev = 0, sms = 0;
  SRK_CUDA_OK(cudaGetDevice(&dev));
  SRK_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // two rows per SM in flight: 296 rows of a few hundred KB stay inside the 126 MB L2 between the
  // two passes, and 512 threads x 128 B per row keep enough loads in flight for HBM
  const int64_t want = (int64_t)sms * 2;
  const unsigned blocks = (unsigned)(R < want ? R : want);
  cudaStream_t st = (cudaStream_t)stream;
  switch (ns) {
    case 1: slice_rows_max_kernel<1><<<blocks, SLICE_THREADS, 0, st>>>(V, ldv, R, K, zero_diag_offset, planes, ldp, plane_stride, bound_out); break;
    case 2: slice_rows_max_kernel<2><<<blocks, SLICE_THREADS, 0, st>>>(V, ldv, R, K, zero_diag_offset, planes, ldp, plane_stride, bound_out); break;
    case 3: slice_rows_max_kernel<3><<<blocks, SLICE_THREADS, 0, st>>>(V, ldv, R, K, zero_diag_offset, planes, ldp, plane_stride, bound_out); break;
    case 4: slice_rows_max_kernel<4><<<blocks, SLICE_THREADS, 0, st>>>(V, ldv, R, K, zero_diag_offset, planes, ldp, plane_stride, bound_out); break;
    default: return srk::fail(SRK_ERR_INVALID, "invalid argument: %s", "ns must be 1..4");
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_slice_rows_key_f64(const double* V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                                      int ns, const uint32_t* rowmax_hi, uint8_t* planes, int64_t ldp,
                                      int64_t plane_stride, double* bound_out, void* stream) {
  SRK_REQUIRE(V && planes && rowmax_hi, "null pointer");
  SRK_REQUIRE(ldp % 16 == 0 && ldp >= K && ldv >= K, "ldp must be a multiple of 16 and >= K");
  SRK_REQUIRE(((uintptr_t)planes % 16) == 0 && plane_stride % 16 == 0, "planes must be 16-byte aligned");
  if (R == 0 || K == 0) return SRK_OK;
  const int64_t threads = R * ((ldp + 15) / 16);
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (ns) {
    case 1: slice_rows_key_kernel<1><<<blocks, 256, 0, st>>>(V, ldv, R, K, zero_diag_offset, rowmax_hi, planes, ldp, plane_stride, bound_out); break;
    case 2: slice_rows_key_kernel<2><<<blocks, 256, 0, st>>>(V, ldv, R, K, zero_diag_offset, rowmax_hi, planes, ldp, plane_stride, bound_out); break;
    case 3: slice_rows_key_kernel<3><<<blocks, 256, 0, st>>>(V, ldv, R, K, zero_diag_offset, rowmax_hi, planes, ldp, plane_stride, bound_out); break;
    case 4: slice_rows_key_kernel<4><<<blocks, 256, 0, st>>>(V, ldv, R, K, zero_diag_offset, rowmax_hi, planes, ldp, plane_stride, bound_out); break;
    default: return srk::fail(SRK_ERR_INVALID, "invalid argument: %s", "ns must be 1..4");
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_topk_rows(const double* S, int64_t lds, int64_t R, int64_t n, int k, int32_t* idx,
                             double* vals, void* stream) {
  SRK_REQUIRE(S && idx && vals, "null pointer");
  SRK_REQUIRE(k >= 0 && k <= n && lds >= n && n < (1ll << 31), "shape");
  if (R == 0 || k == 0) return SRK_OK;
  topk_rows_kernel<<<(unsigned)R, TOPK_THREADS, 0, (cudaStream_t)stream>>>(S, lds, n, k, idx, vals);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_set_identity_f64(double* S, int64_t lds, int64_t R, int64_t n, int64_t diag_offset,
                                    void* stream) {
  SRK_REQUIRE(S && lds >= n, "shape");
  if (R == 0 || n == 0) return SRK_OK;
  const int64_t t = R * n;
  set_identity_kernel<<<(unsigned)((t + 255) / 256), 256, 0, (cudaStream_t)stream>>>(S, lds, R, n, diag_offset);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}
