// bundle.cu -- "row bundle" ordering of the local CSR for H.v (sm_100a).
//
// Why: the plain warp-per-row H.v is bound by the L1TEX data pipe, not by HBM (profiles/r01_spmv_vec32_hci1e7_ncu_full.txt):
// a 32-lane gather of x touches ~30 distinct 32-byte sectors because one row's columns are spread over the whole vector.
// Neighbouring rows of the alpha-major order (same alpha string, adjacent beta strings) connect to NEIGHBOURING columns,
// so the entries of R consecutive rows, merged in column order, gather from far fewer sectors per 32 entries
// (measured on C2: 28 -> 21 (R=2) -> 16.5 (R=4) -> 12.8 (R=8) sectors, 22 -> 6.9 128-byte lines for R=8).
//
// Layout: a bundle = R consecutive local rows; its entries keep their place in cols/vals (the bundle's segment is
// [rowptr[bR], rowptr[bR+R])) but are re-ordered by (column, row) and the column word carries the row-in-bundle in its
// low 3 bits (stored = column << 3 | r).  The byte count is unchanged (12 B per entry) and the encoding is an in-place
// permutation inside each segment, so it is undone exactly (bundle_decode) whenever a consumer wants plain CSR rows
// (export_upper, import); get_row reads a bundle and filters.  One warp multiplies one bundle and keeps
// R running sums per lane (predicated adds: the kernel has ~80% idle issue slots).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "handle.h"
#include "stream_loads.cuh"

namespace sqmc {

static const int kBShift = 3;       // low bits of the stored column word = row within the bundle
static const int kBCapMax = 16384;  // entries of one bundle staged in shared memory by encode/decode (12 B each)

static inline unsigned bblocks(int64_t n, int t = 256) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(div_up(n, t), 0x7fffffff)); }

__global__ void bundle_len_kernel(const int64_t *rowptr, int64_t nloc, int R, int64_t nb, int64_t *len) {
  int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int64_t r1 = min(b * R + R, nloc);
  len[b] = rowptr[r1] - rowptr[b * R];
}

// ------------------------------------------------------------------ encode: plain rows -> (column,row)-ordered bundle
template <int R>
__global__ void __launch_bounds__(512) bundle_encode_kernel(const int64_t *__restrict__ rowptr, int64_t nloc, int64_t nb, int32_t *cols, double *vals, int cap,
                                                            int len_lo, bool take_longer) {
  extern __shared__ __align__(8) unsigned char bsm[];
  double *sv = reinterpret_cast<double *>(bsm);
  int32_t *sc = reinterpret_cast<int32_t *>(sv + cap);
  __shared__ int64_t rp[R + 1];
  for (int64_t b = blockIdx.x; b < nb; b += gridDim.x) {
    if (threadIdx.x <= R) rp[threadIdx.x] = rowptr[min(b * R + (int64_t)threadIdx.x, nloc)];
    __syncthreads();
    const int64_t base = rp[0];
    const int L = (int)min(rp[R] - base, (int64_t)0x7fffffff);
    // this launch handles bundles with len_lo < length <= cap (and the longer, tag-only ones when take_longer):
    // short bundles run with a small shared-memory footprint, i.e. several CTAs per SM
    if (rp[R] - base <= len_lo || (rp[R] - base > cap && !take_longer)) { __syncthreads(); continue; }
    int off[R + 1];
#pragma unroll
    for (int r = 0; r <= R; r++) off[r] = (int)(rp[r] - base);
    if (rp[R] - base > cap) {  // too long to stage: tag only, rows stay contiguous
      for (int64_t i = threadIdx.x; i < rp[R] - base; i += blockDim.x) {
        int r = 0;
#pragma unroll
        for (int q = 1; q < R; q++) r += (base + i >= rp[q]);
        cols[base + i] = (cols[base + i] << kBShift) | r;
      }
    } else {
      for (int i = threadIdx.x; i < L; i += blockDim.x) { sc[i] = cols[base + i]; sv[i] = vals[base + i]; }
      __syncthreads();
      for (int i = threadIdx.x; i < L; i += blockDim.x) {
        int r = 0;
#pragma unroll
        for (int q = 1; q < R; q++) r += (i >= off[q]);
        const int32_t c = sc[i];
        int pos = 0;
#pragma unroll
        for (int q = 0; q < R; q++) {
          if (q == r) { pos += i - off[q]; continue; }
          // entries of row q that sort before (c, r): columns < c, plus an equal column when q < r
          const int32_t key = q < r ? c + 1 : c;
          int lo = off[q], hi = off[q + 1];
          while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (sc[mid] < key) lo = mid + 1;
            else hi = mid;
          }
          pos += lo - off[q];
        }
        cols[base + pos] = (c << kBShift) | r;
        vals[base + pos] = sv[i];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ decode: exact inverse (stable partition by row tag)
template <int R>
__global__ void __launch_bounds__(256) bundle_decode_kernel(const int64_t *__restrict__ rowptr, int64_t nloc, int64_t nb, int32_t *cols, double *vals, int cap) {
  extern __shared__ __align__(8) unsigned char bsm[];
  double *sv = reinterpret_cast<double *>(bsm);
  int32_t *sc = reinterpret_cast<int32_t *>(sv + cap);
  __shared__ int64_t rp[R + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t b = blockIdx.x; b < nb; b += gridDim.x) {
    if (threadIdx.x <= R) rp[threadIdx.x] = rowptr[min(b * R + (int64_t)threadIdx.x, nloc)];
    __syncthreads();
    const int64_t base = rp[0];
    if (rp[R] - base > cap) {
      for (int64_t i = threadIdx.x; i < rp[R] - base; i += blockDim.x) cols[base + i] = cols[base + i] >> kBShift;
    } else {
      const int L = (int)(rp[R] - base);
      for (int i = threadIdx.x; i < L; i += blockDim.x) { sc[i] = cols[base + i]; sv[i] = vals[base + i]; }
      __syncthreads();
      for (int r = w; r < R; r += nw) {
        int64_t out = rp[r];
        for (int i0 = 0; i0 < L; i0 += 32) {
          const int i = i0 + lane;
          const int32_t word = i < L ? sc[i] : -1;
          const bool m = i < L && (word & ((1 << kBShift) - 1)) == r;
          const unsigned bal = __ballot_sync(0xffffffffu, m);
          if (m) {
            const int64_t p = out + __popc(bal & ((1u << lane) - 1));
            cols[p] = word >> kBShift;
            vals[p] = sv[i];
          }
          out += __popc(bal);
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ H.v on bundles: one warp per bundle, R sums per lane
// branch-free on purpose: written as `if (f == r) acc[r] += p` the compiler emits a jump table and the warp diverges
template <int R, int Q>
__device__ __forceinline__ void badd_one(double (&acc)[R], int f, double p) {
  if constexpr (Q < R) {
    asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %2, %3;\n\t@q add.rn.f64 %0, %0, %1;\n\t}" : "+d"(acc[Q]) : "d"(p), "r"(f), "n"(Q));
    badd_one<R, Q + 1>(acc, f, p);
  }
}
template <int R>
__device__ __forceinline__ void badd(double (&acc)[R], int32_t word, double p) {
  badd_one<R, 0>(acc, word & ((1 << kBShift) - 1), p);
}

template <int R>
__global__ void __launch_bounds__(256) spmv_bundle_kernel(const int64_t *__restrict__ rowptr, int64_t nloc, int64_t nb, const int32_t *__restrict__ cols,
                                                          const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y) {
  const Policies P;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nb) return;
  const int64_t r0 = b * R;
  const int64_t e = rowptr[min(r0 + R, nloc)];
  int64_t k = rowptr[r0] + lane;
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) acc[r] = 0.0;
  for (; k + 96 < e; k += 128) {
    const int32_t c0 = ld_col(cols + k, P), c1 = ld_col(cols + k + 32, P), c2 = ld_col(cols + k + 64, P), c3 = ld_col(cols + k + 96, P);
    const double v0 = ld_val(vals + k, P), v1 = ld_val(vals + k + 32, P), v2 = ld_val(vals + k + 64, P), v3 = ld_val(vals + k + 96, P);
    const double x0 = ld_x(x + (c0 >> kBShift), P), x1 = ld_x(x + (c1 >> kBShift), P), x2 = ld_x(x + (c2 >> kBShift), P), x3 = ld_x(x + (c3 >> kBShift), P);
    badd<R>(acc, c0, v0 * x0);
    badd<R>(acc, c1, v1 * x1);
    badd<R>(acc, c2, v2 * x2);
    badd<R>(acc, c3, v3 * x3);
  }
  for (; k < e; k += 32) {
    const int32_t c = ld_col(cols + k, P);
    badd<R>(acc, c, ld_val(vals + k, P) * ld_x(x + (c >> kBShift), P));
  }
  double mine = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    double a = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == r) mine = a;
  }
  if (lane < R && r0 + lane < nloc) y[r0 + lane] = mine;
}

// software-pipelined variant: the column/value loads of block i+1 are issued while the gathers of block i are in flight,
// so a warp always has a block of the 12 B/entry streams outstanding (the plain loop alternates stream and gather phases
// and, once the gathers coalesce, runs out of memory-level parallelism rather than L1TEX throughput)
template <int R, int U>
__global__ void __launch_bounds__(256) spmv_bundle_pipe_kernel(const int64_t *__restrict__ rowptr, int64_t nloc, int64_t nb, const int32_t *__restrict__ cols,
                                                               const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y) {
  const Policies P;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nb) return;
  const int64_t r0 = b * R;
  const int64_t e = rowptr[min(r0 + R, nloc)];
  int64_t kb = rowptr[r0];  // warp-uniform block start
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) acc[r] = 0.0;
  int32_t c[U];
  double v[U];
  bool have = kb + 32 * U <= e;
  if (have) {
#pragma unroll
    for (int u = 0; u < U; u++) c[u] = ld_col(cols + kb + lane + 32 * u, P);
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = ld_val(vals + kb + lane + 32 * u, P);
  }
  while (have) {
    double xx[U];
#pragma unroll
    for (int u = 0; u < U; u++) xx[u] = ld_x(x + (c[u] >> kBShift), P);
    const int64_t kn = kb + 32 * U;
    const bool hn = kn + 32 * U <= e;
    int32_t cn[U];
    double vn[U];
    if (hn) {
#pragma unroll
      for (int u = 0; u < U; u++) cn[u] = ld_col(cols + kn + lane + 32 * u, P);
#pragma unroll
      for (int u = 0; u < U; u++) vn[u] = ld_val(vals + kn + lane + 32 * u, P);
    }
#pragma unroll
    for (int u = 0; u < U; u++) badd<R>(acc, c[u], v[u] * xx[u]);
#pragma unroll
    for (int u = 0; u < U; u++) { c[u] = cn[u]; v[u] = vn[u]; }
    kb = kn;
    have = hn;
  }
  for (int64_t k = kb + lane; k < e; k += 32) {
    const int32_t cc = ld_col(cols + k, P);
    badd<R>(acc, cc, ld_val(vals + k, P) * ld_x(x + (cc >> kBShift), P));
  }
  double mine = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    double a = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == r) mine = a;
  }
  if (lane < R && r0 + lane < nloc) y[r0 + lane] = mine;
}

// two right-hand sides at once (Davidson with n_states >= 2, SURVEY.md 8(d) "SpMM"): x2 / y2 hold the two vectors
// interleaved (x2[2*col + k]), so one 16-byte gather serves both and the matrix is streamed once:
// algorithmic bytes 12*nnz_full + 36*n for two vectors instead of 2*(12*nnz_full + 20*n)
__device__ __forceinline__ double2 bld_x2(const double *p, const Policies &P) {
  double2 r;
  asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(P.x));
  return r;
}
template <int R, int U>
__global__ void __launch_bounds__(256) spmm2_bundle_pipe_kernel(const int64_t *__restrict__ rowptr, int64_t nloc, int64_t nb, const int32_t *__restrict__ cols,
                                                                const double *__restrict__ vals, const double *__restrict__ x2, double *__restrict__ y2) {
  const Policies P;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nb) return;
  const int64_t r0 = b * R;
  const int64_t e = rowptr[min(r0 + R, nloc)];
  int64_t kb = rowptr[r0];
  double acc0[R], acc1[R];
#pragma unroll
  for (int r = 0; r < R; r++) { acc0[r] = 0.0; acc1[r] = 0.0; }
  int32_t c[U];
  double v[U];
  bool have = kb + 32 * U <= e;
  if (have) {
#pragma unroll
    for (int u = 0; u < U; u++) c[u] = ld_col(cols + kb + lane + 32 * u, P);
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = ld_val(vals + kb + lane + 32 * u, P);
  }
  while (have) {
    double2 xx[U];
#pragma unroll
    for (int u = 0; u < U; u++) xx[u] = bld_x2(x2 + 2 * (int64_t)(c[u] >> kBShift), P);
    const int64_t kn = kb + 32 * U;
    const bool hn = kn + 32 * U <= e;
    int32_t cn[U];
    double vn[U];
    if (hn) {
#pragma unroll
      for (int u = 0; u < U; u++) cn[u] = ld_col(cols + kn + lane + 32 * u, P);
#pragma unroll
      for (int u = 0; u < U; u++) vn[u] = ld_val(vals + kn + lane + 32 * u, P);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      badd<R>(acc0, c[u], v[u] * xx[u].x);
      badd<R>(acc1, c[u], v[u] * xx[u].y);
    }
#pragma unroll
    for (int u = 0; u < U; u++) { c[u] = cn[u]; v[u] = vn[u]; }
    kb = kn;
    have = hn;
  }
  for (int64_t k = kb + lane; k < e; k += 32) {
    const int32_t cc = ld_col(cols + k, P);
    const double vv = ld_val(vals + k, P);
    const double2 xv = bld_x2(x2 + 2 * (int64_t)(cc >> kBShift), P);
    badd<R>(acc0, cc, vv * xv.x);
    badd<R>(acc1, cc, vv * xv.y);
  }
  double m0 = 0.0, m1 = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    double a0 = acc0[r], a1 = acc1[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (lane == r) { m0 = a0; m1 = a1; }
  }
  if (lane < R && r0 + lane < nloc) {
    y2[2 * (r0 + lane)] = m0;
    y2[2 * (r0 + lane) + 1] = m1;
  }
}

// ------------------------------------------------------------------ host side
// default: bundles of 4 rows (measured best on B200, profiles/r01_bundle_experiment.txt); SQMC_BUNDLE=0|2|4|8 overrides
static int bundle_want() {
  static int want = -1;
  if (want < 0) {
    const char *e = getenv("SQMC_BUNDLE");
    int v = e ? atoi(e) : 4;
    want = (v == 2 || v == 4 || v == 8) ? v : 0;
  }
  return want;
}
// blocks of 32 entries per lane kept in flight by the pipelined kernel (SQMC_BUNDLE_PIPE=0|2|4, default 4)
static int bundle_pipe() {  // read per launch (a getenv is ~100 ns) so that scripts/bundle_inproc.py can A/B on one resident matrix
  const char *e = getenv("SQMC_BUNDLE_PIPE");
  const int pipe = e ? atoi(e) : 4;
  return (pipe == 0 || pipe == 2 || pipe == 4) ? pipe : 4;
}

template <int R>
static int encode_t(sqmc_b200_handle *h, int cap, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0, nb = div_up(nloc, (int64_t)R);
  SQ_CUDA(cudaFuncSetAttribute(bundle_encode_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 12));
  // two passes: bundles up to kSmall entries with 4 CTAs per SM, the rest with the full staging capacity
  const int kSmall = 4096;
  int lo = 0;
  if (cap > kSmall) {
    bundle_encode_kernel<R><<<(unsigned)std::min<int64_t>(nb, (int64_t)G.sm_count * 4), 512, kSmall * 12, s>>>(h->d_rowptr, nloc, nb, h->d_cols, h->d_vals, kSmall, 0, false);
    SQ_LAUNCH_CHECK();
    lo = kSmall;
  }
  const int smem = cap * 12;
  const int per_sm = std::max(1, std::min(4, (220 * 1024) / std::max(smem, 1)));
  bundle_encode_kernel<R><<<(unsigned)std::min<int64_t>(nb, (int64_t)G.sm_count * per_sm), 512, smem, s>>>(h->d_rowptr, nloc, nb, h->d_cols, h->d_vals, cap, lo, true);
  SQ_LAUNCH_CHECK();
  return 0;
}
template <int R>
static int decode_t(sqmc_b200_handle *h, int cap, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0, nb = div_up(nloc, (int64_t)R);
  const int smem = cap * 12;
  SQ_CUDA(cudaFuncSetAttribute(bundle_decode_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int per_sm = std::max(1, std::min(8, (220 * 1024) / std::max(smem, 1)));
  bundle_decode_kernel<R><<<(unsigned)std::min<int64_t>(nb, (int64_t)G.sm_count * per_sm), 256, smem, s>>>(h->d_rowptr, nloc, nb, h->d_cols, h->d_vals, cap);
  SQ_LAUNCH_CHECK();
  return 0;
}

// plain CSR -> bundles (no-op when disabled, already bundled or the column word has no room for the tag)
int bundle_encode(sqmc_b200_handle *h) { return bundle_encode_r(h, bundle_want()); }
int bundle_encode_r(sqmc_b200_handle *h, int R) {
  const int64_t nloc = h->row1 - h->row0;
  if (!R || h->bundle_R || !h->d_rowptr) return 0;
  if (h->n >= (1ll << (31 - kBShift))) return 0;
  if (nloc == 0 || h->nnz_local == 0) {
    // a rank without rows still switches layout: every rank must take the same (collective) code path in the two-vector H.v
    h->bundle_R = R;
    h->bundle_cap = 1024;
    return 0;
  }
  cudaStream_t s = G.stream;
  if (!h->d_diag) {  // Davidson's preconditioner and the projector read the diagonal: keep a copy
    SQ_CUDA(cudaMalloc(&h->d_diag, nloc * sizeof(double)));
    double *dst = h->d_diag;
    h->d_diag = nullptr;  // extract_diag copies from d_diag when it is set
    int rc = extract_diag(h, dst, s);
    h->d_diag = dst;
    if (rc) return rc;
  }
  // staging capacity = longest bundle (rounded up), capped
  const int64_t nb = div_up(nloc, (int64_t)R);
  DevBuf<int64_t> len, mx;
  SQ_CHECK(len.alloc(nb));
  SQ_CHECK(mx.alloc(1));
  bundle_len_kernel<<<bblocks(nb), 256, 0, s>>>(h->d_rowptr, nloc, R, nb, len.p);
  SQ_LAUNCH_CHECK();
  size_t tb = 0;
  cub::DeviceReduce::Max(nullptr, tb, len.p, mx.p, (int)std::min<int64_t>(nb, 0x7fffffff), s);
  DevBuf<unsigned char> tmp;
  SQ_CHECK(tmp.alloc((int64_t)tb + 16));
  SQ_CUDA(cub::DeviceReduce::Max(tmp.p, tb, len.p, mx.p, (int)std::min<int64_t>(nb, 0x7fffffff), s));
  g_launch_count += 1;
  int64_t maxlen = 0;
  SQ_CUDA(cudaMemcpyAsync(&maxlen, mx.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  SQ_CUDA(cudaStreamSynchronize(s));
  int cap = (int)std::min<int64_t>(kBCapMax, std::max<int64_t>(1024, (maxlen + 1023) / 1024 * 1024));
  if (const char *ce = getenv("SQMC_BUNDLE_CAP")) cap = std::max(64, std::min(cap, atoi(ce)));  // test hook: force the tag-only path
  h->bundle_cap = cap;
  int rc = R == 2 ? encode_t<2>(h, cap, s) : R == 4 ? encode_t<4>(h, cap, s) : encode_t<8>(h, cap, s);
  if (rc) return rc;
  h->bundle_R = R;
  return 0;
}

// bundles -> plain CSR (exact inverse)
int bundle_decode(sqmc_b200_handle *h) {
  if (!h->bundle_R) return 0;
  cudaStream_t s = G.stream;
  const int R = h->bundle_R, cap = h->bundle_cap;
  if (h->row1 - h->row0 > 0 && h->nnz_local > 0) {
    int rc = R == 2 ? decode_t<2>(h, cap, s) : R == 4 ? decode_t<4>(h, cap, s) : decode_t<8>(h, cap, s);
    if (rc) return rc;
  }
  SQ_CUDA(cudaStreamSynchronize(s));
  h->bundle_R = 0;
  return 0;
}

int bundle_spmv(sqmc_b200_handle *h, const double *x, double *y, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0;
  const int R = h->bundle_R;
  const int64_t nb = div_up(nloc, (int64_t)R);
  if (nb == 0) return 0;
  const unsigned grid = (unsigned)div_up(nb * 32, (int64_t)256);
  const int pipe = bundle_pipe();
#define SQ_BL(KERN) KERN<<<grid, 256, 0, s>>>(h->d_rowptr, nloc, nb, h->d_cols, h->d_vals, x, y)
  if (pipe == 4) {
    if (R == 2) SQ_BL((spmv_bundle_pipe_kernel<2, 4>)); else if (R == 4) SQ_BL((spmv_bundle_pipe_kernel<4, 4>)); else SQ_BL((spmv_bundle_pipe_kernel<8, 4>));
  } else if (pipe == 2) {
    if (R == 2) SQ_BL((spmv_bundle_pipe_kernel<2, 2>)); else if (R == 4) SQ_BL((spmv_bundle_pipe_kernel<4, 2>)); else SQ_BL((spmv_bundle_pipe_kernel<8, 2>));
  } else {
    if (R == 2) SQ_BL(spmv_bundle_kernel<2>); else if (R == 4) SQ_BL(spmv_bundle_kernel<4>); else SQ_BL(spmv_bundle_kernel<8>);
  }
#undef SQ_BL
  SQ_LAUNCH_CHECK();
  return 0;
}

// y2 = H x2 for two interleaved vectors; only on bundled matrices (the caller falls back to two H.v otherwise)
int bundle_spmm2(sqmc_b200_handle *h, const double *x2, double *y2, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0;
  const int R = h->bundle_R;
  if (!R) { set_error("bundle_spmm2: matrix is not in row-bundle order"); return 2; }
  const int64_t nb = div_up(nloc, (int64_t)R);
  if (nb == 0) return 0;
  const unsigned grid = (unsigned)div_up(nb * 32, (int64_t)256);
  const char *ue = getenv("SQMC_SPMM_PIPE");
  const int U = ue ? atoi(ue) : 4;  // measured at 10^7 dets: 25.7 ms per pair with 4 blocks in flight, 29.8 ms with 2
#define SQ_BM(KERN) KERN<<<grid, 256, 0, s>>>(h->d_rowptr, nloc, nb, h->d_cols, h->d_vals, x2, y2)
  if (U == 4) {
    if (R == 2) SQ_BM((spmm2_bundle_pipe_kernel<2, 4>)); else if (R == 4) SQ_BM((spmm2_bundle_pipe_kernel<4, 4>)); else SQ_BM((spmm2_bundle_pipe_kernel<8, 2>));
  } else {
    if (R == 2) SQ_BM((spmm2_bundle_pipe_kernel<2, 2>)); else if (R == 4) SQ_BM((spmm2_bundle_pipe_kernel<4, 2>)); else SQ_BM((spmm2_bundle_pipe_kernel<8, 2>));
  }
#undef SQ_BM
  SQ_LAUNCH_CHECK();
  return 0;
}

// one row out of a bundled matrix (get_row): read the bundle's segment and keep the entries tagged with this row
int bundle_get_row(sqmc_b200_handle *h, int64_t internal_row, std::vector<int32_t> &cols, std::vector<double> &vals) {
  const int R = h->bundle_R;
  const int64_t nloc = h->row1 - h->row0, q = internal_row - h->row0;
  const int64_t b = q / R, r = q % R;
  const int64_t ra = b * R, rb = std::min(ra + R, nloc);
  std::vector<int64_t> rp(rb - ra + 1);
  SQ_CUDA(cudaMemcpy(rp.data(), h->d_rowptr + ra, rp.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  const int64_t L = rp.back() - rp[0];
  std::vector<int32_t> c(L);
  std::vector<double> v(L);
  if (L) {
    SQ_CUDA(cudaMemcpy(c.data(), h->d_cols + rp[0], L * sizeof(int32_t), cudaMemcpyDeviceToHost));
    SQ_CUDA(cudaMemcpy(v.data(), h->d_vals + rp[0], L * sizeof(double), cudaMemcpyDeviceToHost));
  }
  cols.clear();
  vals.clear();
  for (int64_t k = 0; k < L; k++)
    if ((c[k] & ((1 << kBShift) - 1)) == r) { cols.push_back(c[k] >> kBShift); vals.push_back(v[k]); }
  return 0;
}

}  // namespace sqmc
