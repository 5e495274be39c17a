// Shared helpers for libsimrank_b200 (error reporting, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "simrank_b200.h"

namespace srk {

// Thread-local error text returned by srk_last_error().
char* error_buffer();

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(error_buffer(), 512, fmt, a, b, c);
  return code;
}

#define SRK_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return srk::fail(SRK_ERR_CUDA, "%s (CUDA error %lld at line %lld)",                   \
                       cudaGetErrorString(e__), (long long)e__, (long long)__LINE__);       \
  } while (0)

#define SRK_REQUIRE(cond, msg)                                                              \
  do {                                                                                      \
    if (!(cond)) return srk::fail(SRK_ERR_INVALID, "invalid argument: %s", msg);            \
  } while (0)

// max over non-negative doubles with an integer atomic: for x >= 0 the IEEE-754 bit pattern is
// monotone in x.  NaN is filtered by the callers.
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr),
            static_cast<unsigned long long>(__double_as_longlong(v)));
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 1 - 0.5^count  (SimRank.py:316).  Exact: 2^-c is a power of two, the subtraction rounds once,
// and the result is exactly 1.0 from c = 54 on, as in numpy.
__device__ __forceinline__ double evidence_factor(unsigned c) {
  if (c >= 54u) return 1.0;
  return 1.0 - __longlong_as_double(static_cast<long long>(1023u - c) << 52);
}

// bound(r) = vec[r]*mul + add  (srk_rowbound)
__device__ __forceinline__ double row_bound(const srk_rowbound& b, int64_t r) {
  return b.vec ? b.vec[r] * b.mul + b.add : b.add;
}

// The epilogue of srk_epilogue applied to one element.  `is_diag` selects fill_diagonal.
struct EpilogueDev {
  double coef, lambda;
  const uint8_t* evidence; int64_t ld_evidence;
  const double* prior; int64_t ld_prior;
  const double* s_old; int64_t ld_s_old;
  int64_t diag_offset;         // CSR half-product: global index of output row 0
};

inline EpilogueDev to_dev(const srk_epilogue& e) {
  EpilogueDev d;
  d.coef = e.coef; d.lambda = e.lambda;
  d.evidence = e.evidence; d.ld_evidence = e.ld_evidence;
  d.prior = e.prior; d.ld_prior = e.ld_prior;
  d.s_old = e.s_old; d.ld_s_old = e.ld_s_old;
  d.diag_offset = e.diag_offset;
  return d;
}

}  // namespace srk
