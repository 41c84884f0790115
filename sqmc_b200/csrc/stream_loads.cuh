// stream_loads.cuh -- cache-policy loads shared by the H.v kernels (spmv.cu, bundle.cu).
#pragma once
#include <cstdint>

namespace sqmc {

// Cache policy (measured on B200, profiles/r01_spmv_variants.txt): the CSR streams are read once ->
// no L1 allocation + L2 evict-first; x is gathered with reuse -> L1 evict-last + L2 evict-last, so the
// 8*n bytes of x stay resident in the 126 MB L2 while 12*nnz bytes stream past it.
struct Policies {
  uint64_t stream, x;
  __device__ __forceinline__ Policies() {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(x));
  }
};
__device__ __forceinline__ int32_t ld_col(const int32_t *p, const Policies &P) {
  int32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(P.stream));
  return r;
}
__device__ __forceinline__ double ld_val(const double *p, const Policies &P) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(P.stream));
  return r;
}
__device__ __forceinline__ double ld_x(const double *p, const Policies &P) {
  double r;
  asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(P.x));
  return r;
}


}  // namespace sqmc
