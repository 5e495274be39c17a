// Microbenchmark: per-SM issue rate of FP64-pipe instructions on this part (DFMA, DMUL, DSETP,
// I2F.F64, F2I.F64) against FP32 FFMA and integer IMAD, with 1..8 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, long long* clk, int iters) {
  double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
  long long i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
  unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3, u4 = 5, u5 = 6, u6 = 7, u7 = 8;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
    if (OP == 1) { f0 = fmaf(f0, 1.0001f, 1e-3f); f1 = fmaf(f1, 1.0001f, 1e-3f); f2 = fmaf(f2, 1.0001f, 1e-3f); f3 = fmaf(f3, 1.0001f, 1e-3f); f4 = fmaf(f4, 1.0001f, 1e-3f); f5 = fmaf(f5, 1.0001f, 1e-3f); f6 = fmaf(f6, 1.0001f, 1e-3f); f7 = fmaf(f7, 1.0001f, 1e-3f); }
    if (OP == 2) { a0 = (double)i0 + a0; i0 += 3; a1 = (double)i1 + a1; i1 += 3; a2 = (double)i2 + a2; i2 += 3; a3 = (double)i3 + a3; i3 += 3; }   // I2F.F64.S64 + DADD
    if (OP == 3) { u0 = u0 * u4 + u1; u1 = u1 * u5 + u2; u2 = u2 * u6 + u3; u3 = u3 * u7 + u0; u4 = u4 * u0 + u5; u5 = u5 * u1 + u6; u6 = u6 * u2 + u7; u7 = u7 * u3 + u4; }
    if (OP == 4) { i0 += (long long)rint(a0); a0 += 1.5; i1 += (long long)rint(a1); a1 += 1.5; i2 += (long long)rint(a2); a2 += 1.5; i3 += (long long)rint(a3); a3 += 1.5; }   // FRND + F2I + DADD
    if (OP == 5) { a0 = a0 > a1 ? a0 * b : a0; a1 = a1 > a2 ? a1 * b : a1; a2 = a2 > a3 ? a2 * b : a2; a3 = a3 > a0 ? a3 * b : a3; }  // DSETP + DMUL
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + i0 + i1 + i2 + i3 + u0 + u1 + u2 + u3 + u4 + u5 + u6 + u7;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int ops_per_iter) {
  double* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
  for (int threads : {128, 256, 512, 1024}) {
    int iters = 4096;
    k<OP><<<148, threads>>>(out, clk, 16);
    cudaDeviceSynchronize();
    k<OP><<<148, threads>>>(out, clk, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = (double)h[0];
    double warp_instrs = (double)iters * ops_per_iter * (threads / 32);
    printf("%-28s threads/SM %4d: %8.2f clk per warp-instr per SM (%.2f lanes/clk/SM)\n", name, threads, cyc / warp_instrs, 32.0 * warp_instrs / cyc);
  }
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<0>("DFMA", 8);
  run<1>("FFMA", 8);
  run<2>("I2F.F64.S64+DADD (pairs)", 4);
  run<3>("IMAD", 8);
  run<4>("FRND+F2I.S64+DADD (triples)", 4);
  run<5>("DSETP+DMUL (pairs)", 4);
  return 0;
}
