// Peak rate of tcgen05.mma kind::i8 (u8 x u8 -> s32) on this GPU: the denominator of
// roofline.frac_int8 in bench.py.  Operands sit in shared memory for the whole run (no TMA, no global
// traffic), one elected thread per CTA (or per CTA pair) issues back-to-back MMAs of the shapes the
// product kernel uses into a TMEM accumulator; the run ends with one commit + wait.
//
//   cta_group::1  M = 128, N = 256, K = 32 per instruction (one SM)
//   cta_group::2  M = 256, N = 256, K = 32 per instruction (two SMs, each holding half of B)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_peak i8_peak.cu
//   ./i8_peak      -> dense int8 Top/s for both groupings, at the clock the run sustained
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {      // K-major, 128 B rows, 128B swizzle, sm_100 descriptor
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int GROUP>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int N = 256, M = GROUP == 2 ? 256 : 128;
  constexpr int b_rows = GROUP == 2 ? N / 2 : N;
  uint8_t* sa = smem;                       // 128 rows x 128 B
  uint8_t* sb = smem + 128 * 128;           // b_rows x 128 B
  for (int i = threadIdx.x; i < (128 + b_rows) * 128; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 1);
  const uint32_t cta = GROUP == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    if (GROUP == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (GROUP == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t acc = tmem_slot;
  constexpr uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D s32, A = B = u8
  unsigned long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0 && cta == 0) {
    const uint64_t da = make_desc(smem_u32(sa)), db = make_desc(smem_u32(sb));
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t accumulate = (it | k) ? 1u : 0u;
        if (GROUP == 2)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(acc), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"(accumulate) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(acc), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"(accumulate) : "memory");
      }
    }
    if (GROUP == 2)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (GROUP == 2) cluster_sync();
  if (threadIdx.x < 32) {
    if (GROUP == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(acc), "r"(256) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(acc), "r"(256) : "memory");
  }
}

template <int GROUP>
static void run(int sms, int iters) {
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, sizeof(unsigned long long) * sms));
  CK(cudaMemset(cyc, 0, sizeof(unsigned long long) * sms));
  const int smem = (128 + 256) * 128 + 1024;
  CK(cudaFuncSetAttribute(peak_kernel<GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = GROUP; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  const int ctas = sms / GROUP * GROUP;
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    CK(cudaLaunchKernelEx(&cfg, peak_kernel<GROUP>, iters, cyc));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double m = GROUP == 2 ? 256.0 : 128.0;
    const double ops = 2.0 * m * 256.0 * 32.0 * 4.0 * iters * (ctas / GROUP);
    unsigned long long h[256];
    CK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost));
    const double clk_per_mma = (double)h[0] / (4.0 * iters);
    printf("cta_group::%d  %d instruction issuers  %8.3f ms  %8.1f Top/s dense int8   %.2f SM clocks per MMA (%.0f op/clk/SM)\n",
           GROUP, ctas / GROUP, ms, ops / ms / 1e9, clk_per_mma, 2.0 * m * 256.0 * 32.0 / clk_per_mma / GROUP);
  }
  CK(cudaFree(cyc));
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("# tcgen05.mma kind::i8 peak, operands resident in shared memory, %d SMs\n", sms);
  run<1>(sms, 200000);
  run<2>(sms, 200000);
  return 0;
}
