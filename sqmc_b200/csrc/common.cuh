// common.cuh -- bit-packed determinant strings, error handling, launch accounting.
// Determinant strings are the reference's integer(ik) (types.f90:26): bit k-1 set
// <=> spatial orbital k occupied.  On the device a string is NW 64-bit words
// (NW=1 when norb <= 64 -- every BASELINE config; NW=2 up to the reference's 127).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace sqmc {

extern std::string g_last_error;
extern int64_t g_launch_count;
extern double g_alloc_stall_ms;  // host milliseconds spent inside allocator / memory-mapping calls (api.cu)
void set_error(const char *fmt, ...);

#define SQ_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      sqmc::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)
#define SQ_CHECK(call)            \
  do {                            \
    int r__ = (call);             \
    if (r__ != 0) return r__;     \
  } while (0)
#define SQ_LAUNCH_CHECK()                                                                        \
  do {                                                                                           \
    sqmc::g_launch_count++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess) {                                                                    \
      sqmc::set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

template <int NW>
struct Bits {
  uint64_t w[NW];
};

template <int NW>
__host__ __device__ __forceinline__ Bits<NW> b_zero() {
  Bits<NW> r;
#pragma unroll
  for (int i = 0; i < NW; i++) r.w[i] = 0;
  return r;
}
template <int NW>
__host__ __device__ __forceinline__ bool b_eq(const Bits<NW> &a, const Bits<NW> &b) {
  bool e = true;
#pragma unroll
  for (int i = 0; i < NW; i++) e = e && (a.w[i] == b.w[i]);
  return e;
}
template <int NW>
__host__ __device__ __forceinline__ bool b_is_zero(const Bits<NW> &a) {
  uint64_t o = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) o |= a.w[i];
  return o == 0;
}
// unsigned multiword compare a < b (the reference compares signed 128-bit; bit 127 is never set)
template <int NW>
__host__ __device__ __forceinline__ bool b_lt(const Bits<NW> &a, const Bits<NW> &b) {
#pragma unroll
  for (int i = NW - 1; i >= 0; i--) {
    if (a.w[i] != b.w[i]) return a.w[i] < b.w[i];
  }
  return false;
}
template <int NW>
__host__ __device__ __forceinline__ Bits<NW> b_and(const Bits<NW> &a, const Bits<NW> &b) {
  Bits<NW> r;
#pragma unroll
  for (int i = 0; i < NW; i++) r.w[i] = a.w[i] & b.w[i];
  return r;
}
template <int NW>
__host__ __device__ __forceinline__ Bits<NW> b_andnot(const Bits<NW> &a, const Bits<NW> &b) {  // a & ~b
  Bits<NW> r;
#pragma unroll
  for (int i = 0; i < NW; i++) r.w[i] = a.w[i] & ~b.w[i];
  return r;
}
template <int NW>
__host__ __device__ __forceinline__ Bits<NW> b_xor(const Bits<NW> &a, const Bits<NW> &b) {
  Bits<NW> r;
#pragma unroll
  for (int i = 0; i < NW; i++) r.w[i] = a.w[i] ^ b.w[i];
  return r;
}
template <int NW>
__device__ __forceinline__ int b_popc(const Bits<NW> &a) {
  int c = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) c += __popcll(a.w[i]);
  return c;
}
// popcount of a ^ b without materialising
template <int NW>
__device__ __forceinline__ int b_popc_xor(const Bits<NW> &a, const Bits<NW> &b) {
  int c = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) c += __popcll(a.w[i] ^ b.w[i]);
  return c;
}
// trailz (position of lowest set bit, 0-based); a != 0
template <int NW>
__device__ __forceinline__ int b_ctz(const Bits<NW> &a) {
#pragma unroll
  for (int i = 0; i < NW; i++) {
    if (a.w[i]) return 64 * i + (__ffsll((long long)a.w[i]) - 1);
  }
  return 64 * NW;
}
template <int NW>
__host__ __device__ __forceinline__ bool b_test(const Bits<NW> &a, int k) {
  if constexpr (NW == 1) return (a.w[0] >> k) & 1ull;  // no dynamic word index: keeps the string in registers
  else return ((k < 64 ? a.w[0] : a.w[1]) >> (k & 63)) & 1ull;
}
template <int NW>
__host__ __device__ __forceinline__ void b_clear(Bits<NW> &a, int k) {
  if constexpr (NW == 1) {
    a.w[0] &= ~(1ull << k);
  } else {
    const uint64_t m = ~(1ull << (k & 63));
    if (k < 64) a.w[0] &= m;
    else a.w[1] &= m;
  }
}
template <int NW>
__host__ __device__ __forceinline__ void b_set(Bits<NW> &a, int k) {
  if constexpr (NW == 1) {
    a.w[0] |= (1ull << k);
  } else {
    const uint64_t m = 1ull << (k & 63);
    if (k < 64) a.w[0] |= m;
    else a.w[1] |= m;
  }
}
// clear lowest set bit
template <int NW>
__device__ __forceinline__ void b_clear_lowest(Bits<NW> &a) {
#pragma unroll
  for (int i = 0; i < NW; i++) {
    if (a.w[i]) {
      a.w[i] &= a.w[i] - 1;
      return;
    }
  }
}
// maskr(k): the k lowest bits set (Fortran maskr)
template <int NW>
__device__ __forceinline__ Bits<NW> b_maskr(int k) {
  Bits<NW> r;
#pragma unroll
  for (int i = 0; i < NW; i++) {
    int kk = k - 64 * i;
    r.w[i] = (kk <= 0) ? 0ull : ((kk >= 64) ? ~0ull : ((1ull << kk) - 1ull));
  }
  return r;
}
template <int NW>
__device__ __forceinline__ Bits<NW> b_load(const uint64_t *p, int64_t i) {
  Bits<NW> r;
#pragma unroll
  for (int k = 0; k < NW; k++) r.w[k] = p[i * NW + k];
  return r;
}
template <int NW>
__device__ __forceinline__ void b_store(uint64_t *p, int64_t i, const Bits<NW> &v) {
#pragma unroll
  for (int k = 0; k < NW; k++) p[i * NW + k] = v.w[k];
}

static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace sqmc
