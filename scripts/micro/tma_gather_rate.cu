// Microbenchmark behind the CSR gather design (DESIGN.md "K3"): how fast can one SM pull random
// L2-resident row segments into shared memory, as a function of the mechanism and the segment size?
//
//   bulk     cp.async.bulk (1-D TMA), one segment per instruction
//   gather4  cp.async.bulk.tensor.2d.tile::gather4, four rows per instruction
//   ldg      plain 16-byte loads by the whole warp, U segments in flight per warp (no smem)
//
// Every warp owns a ring of D slots; one lane issues, all lanes wait on the mbarrier and read one
// word of the landed segment (so the data really arrives).  The source is a `rows x rowbytes` matrix
// (default 32768 x 1024 B = 32 MB: L2 resident after the first pass), row indices are pseudo-random.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather_rate tma_gather_rate.cu -lcuda
//   ./tma_gather_rate            prints GB/s and cycles per instruction per SM for each variant
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* map, uint64_t* b, int c, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(b)), "r"(c), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ uint32_t rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// mode 0 bulk, 1 gather4.  seg = bytes per row segment, depth = slots per warp.
template <int MODE>
__global__ void ring_kernel(const uint8_t* src, const __grid_constant__ CUtensorMap map, int rows, int rowbytes, int seg,
                            int depth, int iters, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int slot_bytes = MODE == 1 ? 4 * seg : seg;
  uint8_t* ring = smem + (size_t)warp * depth * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nw * depth * slot_bytes) + warp * depth;
  if (lane < depth) mbar_init(&bars[lane], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t seed = (blockIdx.x * 64 + warp) * 2654435761u + 12345u;
  const int col0 = (blockIdx.x % (rowbytes / seg)) * seg;           // a column panel per CTA, like the real kernel
  unsigned acc = 0;
  auto issue = [&](int slot) {
    if (MODE == 0) {
      const int r = rnd(seed) % rows;
      mbar_expect(&bars[slot], seg);
      bulk(ring + slot * slot_bytes, src + (size_t)r * rowbytes + col0, seg, &bars[slot]);
    } else {
      const int r0 = rnd(seed) % rows, r1 = rnd(seed) % rows, r2 = rnd(seed) % rows, r3 = rnd(seed) % rows;
      mbar_expect(&bars[slot], 4 * seg);
      gather4(ring + slot * slot_bytes, &map, &bars[slot], col0, r0, r1, r2, r3);
    }
  };
  if (lane == 0) for (int s = 0; s < depth; ++s) issue(s);
  for (int it = 0; it < iters; ++it) {
    const int slot = it % depth;
    mbar_wait(&bars[slot], (it / depth) & 1);
    acc += reinterpret_cast<const unsigned*>(ring + slot * slot_bytes)[lane];
    __syncwarp();
    if (lane == 0 && it + depth < iters) issue(slot);
  }
  if (acc == 0xdeadbeef) sink[0] = acc;
}

// plain loads: each lane 16 B, `seg/512` loads per segment, U segments in flight
template <int U>
__global__ void ldg_kernel(const uint8_t* src, int rows, int rowbytes, int seg, int iters, unsigned* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t seed = (blockIdx.x * 64 + warp) * 2654435761u + 12345u;
  const int col0 = (blockIdx.x % (rowbytes / seg)) * seg;
  unsigned acc = 0;
  for (int it = 0; it < iters; it += U) {
    uint4 v[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rnd(seed) % rows;
      const uint4* p = reinterpret_cast<const uint4*>(src + (size_t)r * rowbytes + col0) + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k * 512 < seg) v[u][k] = __ldg(p + 32 * k);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k * 512 < seg) acc += v[u][k].x + v[u][k].w;
  }
  if (acc == 0xdeadbeef) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int rows = 32768, rowbytes = 1024;
  uint8_t* src;
  unsigned* sink;
  CK(cudaMalloc(&src, (size_t)rows * rowbytes));
  CK(cudaMemset(src, 1, (size_t)rows * rowbytes));
  CK(cudaMalloc(&sink, 4));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  printf("# SMs %d, nominal %.0f MHz; source %d x %d B\n", sms, clk_khz / 1e3, rows, rowbytes);
  printf("%-8s %5s %5s %5s %5s %9s %12s %14s\n", "mode", "seg", "warps", "depth", "ctas", "GB/s", "B/clk/SM", "clk/instr/SM");
  const int segs[] = {256, 512, 1024};
  for (int mode = 0; mode < 2; ++mode)
    for (int si = 0; si < 3; ++si)
      for (int warps = 4; warps <= 16; warps *= 2)
        for (int ctas = 1; ctas <= 2; ++ctas) {
          const int seg = segs[si];
          int depth = 8;
          const int slot = mode == 1 ? 4 * seg : seg;
          while ((size_t)warps * depth * slot + warps * depth * 8 > 100 * 1024 && depth > 2) depth /= 2;
          const size_t smem = (size_t)warps * depth * slot + warps * depth * 8;
          CUtensorMap map;
          memset(&map, 0, sizeof(map));
          {
            cuuint64_t dims[2] = {(cuuint64_t)rowbytes, (cuuint64_t)rows};
            cuuint64_t strides[1] = {(cuuint64_t)rowbytes};
            cuuint32_t box[2] = {(cuuint32_t)seg, 1};
            cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d (seg %d)\n", (int)r, seg); continue; }
          }
          const int iters = mode == 1 ? 2048 : 8192;
          auto run = [&]() {
            if (mode == 0) {
              CK(cudaFuncSetAttribute(ring_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              ring_kernel<0><<<sms * ctas, warps * 32, smem>>>(src, map, rows, rowbytes, seg, depth, iters, sink);
            } else {
              CK(cudaFuncSetAttribute(ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              ring_kernel<1><<<sms * ctas, warps * 32, smem>>>(src, map, rows, rowbytes, seg, depth, iters, sink);
            }
          };
          run();
          CK(cudaDeviceSynchronize());
          CK(cudaEventRecord(a));
          run();
          CK(cudaEventRecord(b));
          CK(cudaEventSynchronize(b));
          float ms = 0;
          CK(cudaEventElapsedTime(&ms, a, b));
          const double bytes = (double)sms * ctas * warps * iters * slot;
          const double instr_per_sm = (double)ctas * warps * iters;
          const double clk = ms * 1e-3 * 1.965e9;
          printf("%-8s %5d %5d %5d %5d %9.0f %12.1f %14.1f\n", mode ? "gather4" : "bulk", seg, warps, depth, ctas, bytes / ms / 1e6,
                 bytes / sms / clk, clk / instr_per_sm);
        }
  for (int si = 0; si < 3; ++si)
    for (int warps = 8; warps <= 32; warps *= 2)
      for (int ctas = 1; ctas <= 2; ++ctas) {
        const int seg = segs[si] < 512 ? 512 : segs[si];
        if (si == 0) continue;
        const int iters = 8192;
        ldg_kernel<4><<<sms * ctas, warps * 32>>>(src, rows, rowbytes, seg, iters, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        ldg_kernel<4><<<sms * ctas, warps * 32>>>(src, rows, rowbytes, seg, iters, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        const double bytes = (double)sms * ctas * warps * iters * seg;
        const double clk = ms * 1e-3 * 1.965e9;
        printf("%-8s %5d %5d %5d %5d %9.0f %12.1f %14s\n", "ldg x4", seg, warps, 4, ctas, bytes / ms / 1e6, bytes / sms / clk, "-");
      }
  return 0;
}
