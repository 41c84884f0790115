// wcsr.cu -- window-staged layout of H for dense determinant spaces and its H.v kernel (sm_100a).
//
// Why: in the plain CSR kernel (spmv.cu) the 8-byte gathers of x dominate: every gather pulls a
// 32-byte sector through L1TEX/L2 (profiles/r01_spmv_vec32_1e7_ncu_full.txt: L1TEX 86 %, L2 81 %,
// DRAM 73 %).  The structure of H removes that cost.  With rows in alpha-major order
//   * a row (u,d) couples to its own alpha-group (dn excitations) and to the alpha-groups of the
//     strings one excitation away from u (opposite-spin doubles): ~90 column groups, shared by ALL
//     rows of the alpha-group;
//   * the remaining entries (same dn, up excitations) are scattered in alpha-major order but are
//     confined to the row's own beta-group in beta-major order.
// So H is split into part A (alpha-major rows, columns in own + neighbour alpha-groups) and part B
// (beta-major rows, columns in the own beta-group).  Each part is cut into tiles of <= 256 rows of one
// row group; the entries of a tile are stored window-major -- (column window, row, column) -- as
// packed (row_local << 16 | col_local) + f64: still 12 bytes per entry.  One CTA owns a tile; its 16
// warps each consume work items (one window piece, <= 2048 consecutive entries): the warp brings the
// x slice of its NEXT item into its private shared-memory buffer with a bulk async copy
// (cp.async.bulk + mbarrier, double buffered) while it streams the current item with unrolled
// evict-first loads, gathers x from shared memory, does a shuffle segmented reduction by row and
// accumulates into its private shared-memory row sums; the 16 partial sums of a row are added in a
// fixed order at the end of the tile (deterministic, no atomics, no block barrier in the hot loop).
// y = A x + P^T B (P x), P = alpha-major -> beta-major.
//
// The layout is produced from the CSR the build emits (pure permutation of entries, in place, with
// bounded scratch) and decoded back on the host for export / get_row.  It replaces, for large dense
// spaces, the same reference routine as spmv.cu: fast_sparse_matrix_multiply_upper_triangular
// (more_tools.f90:3622) and its column-band variant (:3562).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "handle.h"

namespace sqmc {

static const int kWThreads = kWWarps * 32;  // 512

static inline unsigned gblocks(int64_t n, int t = 256) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(div_up(n, t), 0x7fffffff)); }

// ------------------------------------------------------------------ small kernels
__global__ void extract_diag_kernel(const int64_t *rowptr, const int32_t *cols, const double *vals, int64_t row0, int64_t nloc, double *diag) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= nloc) return;
  int64_t lo = rowptr[q], hi = rowptr[q + 1];
  int32_t target = (int32_t)(row0 + q);
  double d = 0.0;
  while (lo < hi) {  // columns of a row are ascending
    int64_t mid = (lo + hi) >> 1;
    int32_t c = cols[mid];
    if (c == target) { d = vals[mid]; break; }
    if (c < target) lo = mid + 1;
    else hi = mid;
  }
  diag[q] = d;
}

int extract_diag(sqmc_b200_handle *h, double *diag_dev, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0;
  if (nloc == 0) return 0;
  if (h->d_diag) {
    SQ_CUDA(cudaMemcpyAsync(diag_dev, h->d_diag, nloc * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  extract_diag_kernel<<<gblocks(nloc), 256, 0, s>>>(h->d_rowptr, h->d_cols, h->d_vals, h->row0, nloc, diag_dev);
  SQ_LAUNCH_CHECK();
  return 0;
}

__global__ void colwin_kernel(const int32_t *grp_of_col, const int64_t *goff, const int32_t *wbase, int32_t *colwin, int64_t ncols) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  int32_t g = grp_of_col[c];
  colwin[c] = wbase[g] + (int32_t)((c - goff[g]) / kWinMax);
}

struct InRange {
  int32_t lo, hi;
  __device__ __forceinline__ bool operator()(const int32_t &v) const { return v >= lo && v < hi; }
};
__global__ void invert_local_kernel(const int32_t *browL, int32_t *inv, int64_t row0, int64_t nloc) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nloc) inv[browL[i] - row0] = (int32_t)i;
}
// per local row: number of same-beta (different alpha) entries -> part B
template <int NW>
__global__ void __launch_bounds__(256) classify_kernel(const int64_t *rowptr, const int32_t *cols, const uint64_t *dn, int64_t row0, int64_t nloc, int32_t *cntA,
                                                       int32_t *cntB_brow, const int32_t *browL_inv) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= nloc) return;
  const int64_t p = row0 + r;
  const Bits<NW> dp = b_load<NW>(dn, p);
  int nb = 0;
  const int64_t s = rowptr[r], e = rowptr[r + 1];
  for (int64_t k = s + lane; k < e; k += 32) {
    int32_t c = cols[k];
    if (c != p && b_eq(b_load<NW>(dn, c), dp)) nb++;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(0xffffffffu, nb, o);
  if (lane == 0) {
    cntA[r] = (int32_t)(e - s) - nb;
    cntB_brow[browL_inv[r]] = nb;
  }
}
// split a chunk of rows: A entries compacted into scratch, B entries written to their beta-major rows
template <int NW>
__global__ void __launch_bounds__(256) split_kernel(const int64_t *rowptr, const int32_t *cols, const double *vals, const uint64_t *dn, int64_t row0,
                                                    int64_t r_begin, int64_t r_end, const int64_t *rowptrA, const int64_t *rowptrB,
                                                    const int32_t *browL_inv, const int32_t *binv, int32_t *scrA_cols, double *scrA_vals,
                                                    int32_t *colsB, double *valsB) {
  const int lane = threadIdx.x & 31;
  const int64_t r = r_begin + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (r >= r_end) return;
  const int64_t p = row0 + r;
  const Bits<NW> dp = b_load<NW>(dn, p);
  const int64_t s = rowptr[r], e = rowptr[r + 1];
  int64_t wa = rowptrA[r] - rowptrA[r_begin], wb = rowptrB[browL_inv[r]];
  const unsigned lt = (1u << lane) - 1u;
  for (int64_t kb = s; kb < e; kb += 32) {
    int64_t k = kb + lane;
    bool in = k < e;
    int32_t c = in ? cols[k] : 0;
    double v = in ? vals[k] : 0.0;
    bool isB = in && c != p && b_eq(b_load<NW>(dn, c), dp);
    unsigned mB = __ballot_sync(0xffffffffu, isB), mA = __ballot_sync(0xffffffffu, in && !isB);
    if (isB) {
      int64_t q = wb + __popc(mB & lt);
      colsB[q] = binv[c];
      valsB[q] = v;
    } else if (in) {
      int64_t q = wa + __popc(mA & lt);
      scrA_cols[q] = c;
      scrA_vals[q] = v;
    }
    wa += __popc(mA);
    wb += __popc(mB);
  }
}

// ------------------------------------------------------------------ tile conversion
// rank of a global window id among the windows present in the tile
__device__ __forceinline__ int win_rank(const uint32_t *bitmap, const uint32_t *wprefix, int32_t id) {
  return (int)wprefix[id >> 5] + __popc(bitmap[id >> 5] & ((1u << (id & 31)) - 1u));
}

// pass 1: number of work items per tile (windows present, entries per window).  cntw: per-CTA scratch of
// ngwin counters indexed by global window id (kept zeroed between tiles).
__global__ void __launch_bounds__(kWThreads) tile_count_items_kernel(WPart P, const int32_t *cols, int nwords, int32_t *cntw_all, int64_t *item_count,
                                                                     int *max_windows) {
  extern __shared__ uint32_t sm1[];
  uint32_t *bitmap = sm1;
  __shared__ int total, totalw;
  int32_t *cntw = cntw_all + (int64_t)blockIdx.x * P.ngwin;
  for (int64_t t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) bitmap[i] = 0;
    if (threadIdx.x == 0) { total = 0; totalw = 0; }
    __syncthreads();
    const int64_t e0 = P.tile_ent0[t], e1 = P.tile_ent0[t + 1];
    for (int64_t k = e0 + threadIdx.x; k < e1; k += blockDim.x) {
      int32_t id = P.colwin[cols[P.ent0 + k]];
      atomicOr(&bitmap[id >> 5], 1u << (id & 31));
      atomicAdd(&cntw[id], 1);
    }
    __syncthreads();
    int c = 0, cw = 0;
    for (int i = threadIdx.x; i < nwords * 32; i += blockDim.x) {
      if (bitmap[i >> 5] & (1u << (i & 31))) {
        c += (cntw[i] + kItemMax - 1) / kItemMax;
        cw++;
        cntw[i] = 0;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); cw += __shfl_xor_sync(0xffffffffu, cw, o); }
    if ((threadIdx.x & 31) == 0 && c) { atomicAdd(&total, c); atomicAdd(&totalw, cw); }
    __syncthreads();
    if (threadIdx.x == 0) { item_count[t] = total; atomicMax(max_windows, totalw); }
    __syncthreads();
  }
}

// pass 2: reorder the entries of a tile from (row, col) to (window, row, col), pack indices, write work items.
// cnt: [W][R] counts -> offsets, in shared memory when it fits (cnt_smem_ints), else in per-CTA global scratch.
__global__ void __launch_bounds__(kWThreads) tile_reorder_kernel(WPart P, int32_t *cols, double *vals, int nwords, int cnt_smem_ints,
                                                                 int32_t *cnt_scratch, int64_t cnt_scratch_stride, int32_t *scr_idx,
                                                                 double *scr_val, int64_t scr_stride) {
  extern __shared__ uint32_t sm2[];
  uint32_t *bitmap = sm2;               // nwords
  uint32_t *wprefix = sm2 + nwords;     // nwords
  int32_t *cnt_sm = (int32_t *)(sm2 + 2 * nwords);
  __shared__ int32_t part_sums[kWThreads];
  __shared__ int W_sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t *my_scr_idx = scr_idx + (int64_t)blockIdx.x * scr_stride;
  double *my_scr_val = scr_val + (int64_t)blockIdx.x * scr_stride;
  for (int64_t t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
    const int R = P.tile_nrows[t];
    const int32_t trow0 = P.tile_row0[t];
    const int64_t e0 = P.tile_ent0[t], e1 = P.tile_ent0[t + 1];
    const int64_t E = e1 - e0;
    int32_t *tc = cols + P.ent0 + e0;
    double *tv = vals + P.ent0 + e0;
    // ---- a. windows present
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) bitmap[i] = 0;
    __syncthreads();
    for (int64_t k = threadIdx.x; k < E; k += blockDim.x) {
      int32_t id = P.colwin[tc[k]];
      atomicOr(&bitmap[id >> 5], 1u << (id & 31));
    }
    __syncthreads();
    // ---- b. exclusive prefix of popcounts over the bitmap words
    {
      const int per = (nwords + kWThreads - 1) / kWThreads;
      int s = 0;
      for (int i = threadIdx.x * per; i < min(nwords, (int)(threadIdx.x + 1) * per); i++) s += __popc(bitmap[i]);
      part_sums[threadIdx.x] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < kWThreads; i++) { int v = part_sums[i]; part_sums[i] = run; run += v; }
        W_sh = run;
      }
      __syncthreads();
      int run = part_sums[threadIdx.x];
      for (int i = threadIdx.x * per; i < min(nwords, (int)(threadIdx.x + 1) * per); i++) { wprefix[i] = run; run += __popc(bitmap[i]); }
      __syncthreads();
    }
    const int W = W_sh;
    const int64_t N = (int64_t)W * R;
    int32_t *cnt = (N <= cnt_smem_ints) ? cnt_sm : (cnt_scratch + (int64_t)blockIdx.x * cnt_scratch_stride);
    // ---- c. counts per (window, row)
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int r = warp; r < R; r += kWWarps) {
      const int64_t s = P.rowptr[trow0 + r] - e0, e = P.rowptr[trow0 + r + 1] - e0;
      for (int64_t k = s + lane; k < e; k += 32) {
        int w = win_rank(bitmap, wprefix, P.colwin[tc[k]]);
        atomicAdd(&cnt[(int64_t)w * R + r], 1);
      }
    }
    __syncthreads();
    // ---- d. exclusive scan of cnt in (window-major, row) order
    {
      const int64_t per = (N + kWThreads - 1) / kWThreads;
      const int64_t i0 = threadIdx.x * per, i1 = min(N, i0 + per);
      int s = 0;
      for (int64_t i = i0; i < i1; i++) s += cnt[i];
      part_sums[threadIdx.x] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < kWThreads; i++) { int v = part_sums[i]; part_sums[i] = run; run += v; }
      }
      __syncthreads();
      int run = part_sums[threadIdx.x];
      for (int64_t i = i0; i < i1; i++) { int v = cnt[i]; cnt[i] = run; run += v; }
      __syncthreads();
    }
    // ---- e. work items: each window's entry range cut into pieces of <= kItemMax entries (serial, ~W steps)
    if (threadIdx.x == 0) {
      WItem *it = P.items + P.tile_item0[t];
      int w = 0;
      for (int word = 0; word < nwords; word++) {
        uint32_t bits = bitmap[word];
        while (bits) {
          int id = word * 32 + (__ffs(bits) - 1);
          bits &= bits - 1;
          int32_t kb = cnt[(int64_t)w * R], ke = (w + 1 < W) ? cnt[(int64_t)(w + 1) * R] : (int32_t)E;
          for (int32_t k = kb; k < ke; k += kItemMax) {
            WItem v;
            v.col0 = P.gwin_col0[id];
            v.len = P.gwin_len[id];
            v.k0 = k;
            v.k1 = min(ke, k + kItemMax);
            *it++ = v;
          }
          w++;
        }
      }
    }
    // ---- f. scatter to scratch in (window, row, col) order
    for (int r = warp; r < R; r += kWWarps) {
      const int64_t s = P.rowptr[trow0 + r] - e0, e = P.rowptr[trow0 + r + 1] - e0;
      int carry_w = -1, carry_n = 0;  // run continuing from the previous 32-entry chunk
      for (int64_t kb = s; kb < e; kb += 32) {
        int64_t k = kb + lane;
        bool in = k < e;
        int32_t c = in ? tc[k] : 0;
        int32_t id = in ? P.colwin[c] : 0;
        int w = in ? win_rank(bitmap, wprefix, id) : -2;
        int wprev = __shfl_up_sync(0xffffffffu, w, 1);
        bool head = in && (lane == 0 || w != wprev);
        unsigned hm = __ballot_sync(0xffffffffu, head);
        int last_head = 31 - __clz(hm & ((2u << lane) - 1u));  // lane of the head of my run (lane 0 is always a head)
        int rel = lane - last_head;
        if (last_head == 0 && w == carry_w) rel += carry_n;
        if (in) {
          int64_t dst = (int64_t)cnt[(int64_t)w * R + r] + rel;
          my_scr_idx[dst] = (int32_t)(((uint32_t)r << 16) | (uint32_t)(c - P.gwin_col0[id]));
          my_scr_val[dst] = tv[k];
        }
        int w31 = __shfl_sync(0xffffffffu, w, 31), rel31 = __shfl_sync(0xffffffffu, rel, 31);
        carry_w = w31;
        carry_n = rel31 + 1;
      }
    }
    __syncthreads();
    // ---- g. copy back
    for (int64_t k = threadIdx.x; k < E; k += blockDim.x) {
      tc[k] = my_scr_idx[k];
      tv[k] = my_scr_val[k];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ the WCSR H.v kernel
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct StreamPol {
  uint64_t p;
  __device__ __forceinline__ StreamPol() { asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); }
};
__device__ __forceinline__ uint32_t ld_idx(const int32_t *p, const StreamPol &S) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(S.p));
  return r;
}
__device__ __forceinline__ double ld_v(const double *p, const StreamPol &S) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(S.p));
  return r;
}

__device__ __forceinline__ uint4 ld_idx4(const int32_t *p, const StreamPol &S) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(S.p));
  return r;
}
__device__ __forceinline__ double2 ld_v2(const double *p, const StreamPol &S) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(S.p));
  return r;
}

// 256 consecutive entries, 8 per lane, loaded with 128-bit loads (k is a multiple of 8 in absolute entry index).
// Entries outside [kmin, kmax) (the unaligned head / the tail of an item) are masked: product 0 and the row id of
// the nearest valid entry, so they merge into an existing segment.  Each lane sums its 8 entries sequentially
// (branch-free running segmented sum); rows closed inside a lane are flushed by that lane; the open last
// segments are combined across lanes with ONE segmented scan per 256 entries.
struct VecRegs {
  uint4 a, b;
  double2 v0, v1, v2, v3;
};
__device__ __forceinline__ VecRegs vec_load(const int32_t *pidx, const double *pval, int64_t k, int lane, const StreamPol &SP) {
  VecRegs r;
  const int64_t base = k + lane * 8;
  r.a = ld_idx4(pidx + base, SP);
  r.b = ld_idx4(pidx + base + 4, SP);
  r.v0 = ld_v2(pval + base, SP);
  r.v1 = ld_v2(pval + base + 2, SP);
  r.v2 = ld_v2(pval + base + 4, SP);
  r.v3 = ld_v2(pval + base + 6, SP);
  return r;
}
// k: first entry of the step (may be < kmin); valid entries are [kmin, kmax); row_first/row_last: rows of entries kmin / kmax-1
__device__ __forceinline__ void vec_consume(const VecRegs &g, int64_t k, int64_t kmin, int64_t kmax, int row_first, int row_last, int lane,
                                            const double *xb, double *yacc) {
  const uint32_t u[8] = {g.a.x, g.a.y, g.a.z, g.a.w, g.b.x, g.b.y, g.b.z, g.b.w};
  const double v[8] = {g.v0.x, g.v0.y, g.v1.x, g.v1.y, g.v2.x, g.v2.y, g.v3.x, g.v3.y};
  const int64_t base = k + lane * 8;
  int r[8];
  double p[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int64_t kk = base + j;
    const bool lo = kk < kmin, hi = kk >= kmax;
    r[j] = lo ? row_first : (hi ? row_last : (int)(u[j] >> 16));
    const double xv = xb[u[j] & 0x3ffu];  // col_local <= kWinMax + 1 < 1024; masked entries read a harmless slot
    p[j] = (lo || hi) ? 0.0 : v[j] * xv;
  }
  // running segmented sums inside the lane
  double sum[8];
  sum[0] = p[0];
#pragma unroll
  for (int j = 1; j < 8; j++) sum[j] = (r[j] == r[j - 1]) ? sum[j - 1] + p[j] : p[j];
  // segments closed inside the lane: the first one is kept (it may continue the previous lane's row), later ones are complete rows
  bool multi = false;
  double af = 0.0;
#pragma unroll
  for (int j = 0; j < 7; j++) {
    if (r[j + 1] != r[j]) {
      if (!multi) { af = sum[j]; multi = true; }
      else yacc[r[j]] += sum[j];
    }
  }
  const int rf = r[0], rl = r[7];
  double al = sum[7];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, al, d);
    int kk = __shfl_up_sync(0xffffffffu, rl, d);
    if (lane >= d && kk == rl) al += t;
  }
  const int rl_next = __shfl_down_sync(0xffffffffu, rl, 1);
  const int rl_prev = __shfl_up_sync(0xffffffffu, rl, 1);
  const int rf_next = __shfl_down_sync(0xffffffffu, rf, 1);
  const double af_next = __shfl_down_sync(0xffffffffu, af, 1);
  const int multi_next = __shfl_down_sync(0xffffffffu, (int)multi, 1);
  if (lane == 31 || rl_next != rl) {
    double tot = al;
    if (lane < 31 && multi_next && rf_next == rl) tot += af_next;  // the next lane's first segment continues this row
    yacc[rl] += tot;
  }
  if (multi && (lane == 0 || rl_prev != rf)) yacc[rf] += af;  // first segment that does not continue the previous lane's row
  __syncwarp();
}

static const int kXwStride = kWinMax + 2;  // doubles per staging buffer
static const int kWcsrSmemBytes = kWWarps * 2 * kXwStride * 8 + kWWarps * kTileRows * 8 + kWWarps * 2 * 8;

// Persistent CTAs (one per SM) take tiles from an atomic counter.  x: the part's column space (length ncols);
// y: the part's local rows.
__global__ void __launch_bounds__(kWThreads, 1) wcsr_spmv_kernel(WPart P, const int32_t *__restrict__ idx, const double *__restrict__ vals,
                                                                 const double *__restrict__ x, int64_t ncols, double *__restrict__ y,
                                                                 unsigned long long *tile_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *xw_all = reinterpret_cast<double *>(smem_raw);                                   // [warp][2][kXwStride]
  double *yacc_all = xw_all + kWWarps * 2 * kXwStride;                                     // [warp][kTileRows]
  uint64_t *bars_all = reinterpret_cast<uint64_t *>(yacc_all + kWWarps * kTileRows);       // [warp][2]
  __shared__ long long s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *xw = xw_all + warp * 2 * kXwStride;
  double *yacc = yacc_all + warp * kTileRows;
  uint64_t *bar = bars_all + warp * 2;
  const StreamPol SP;
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase0 = 0, phase1 = 0;
  const int32_t *pidx = idx + P.ent0;
  const double *pval = vals + P.ent0;
  while (true) {
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(tile_counter, 1ull);
    __syncthreads();
    const int64_t t = s_tile;
    if (t >= P.ntiles) break;
    const int R = P.tile_nrows[t];
    const int64_t e0 = P.tile_ent0[t];
    const int64_t i0 = P.tile_item0[t], ni = P.tile_item0[t + 1] - i0;
    for (int i = lane; i < kTileRows; i += 32) yacc[i] = 0.0;
    // stage the x slice of item j into buffer b (lane 0): bulk copy of the in-range even part + (rare) tail
    auto stage = [&](int64_t j, int b) {
      if (lane == 0) {
        const int4 itv = *reinterpret_cast<const int4 *>(P.items + i0 + j);
        const int32_t c0 = itv.x, len = itv.y;
        int32_t blen = len;
        if ((int64_t)c0 + blen > ncols) blen = (int32_t)((ncols - c0) & ~1ll);
        if (blen > 0) {
          mbar_expect_tx(&bar[b], (uint32_t)blen * 8u);
          bulk_g2s(xw + b * kXwStride, x + c0, (uint32_t)blen * 8u, &bar[b]);
        } else {
          mbar_arrive(&bar[b]);
        }
        for (int i = blen; i < len && (int64_t)c0 + i < ncols; i++) xw[b * kXwStride + i] = x[c0 + i];
      }
    };
    if (warp < ni) stage(warp, 0);
    __syncwarp();
    int b = 0;
    for (int64_t j = warp; j < ni; j += kWWarps, b ^= 1) {
      if (j + kWWarps < ni) stage(j + kWWarps, b ^ 1);
      const int4 itv = *reinterpret_cast<const int4 *>(P.items + i0 + j);
      const int64_t k0 = e0 + itv.z, k1 = e0 + itv.w;
      if (b == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
      else { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
      const double *xb = xw + b * kXwStride;
      {
        // masked 256-entry vector steps from the 8-aligned (absolute index) start; 2 steps of loads kept in flight
        const int row_first = (int)((uint32_t)__ldg(pidx + k0) >> 16), row_last = (int)((uint32_t)__ldg(pidx + k1 - 1) >> 16);
        int64_t k = k0 - ((P.ent0 + k0) & 7);
        VecRegs c0 = vec_load(pidx, pval, k, lane, SP);
        VecRegs c1;
        if (k + 256 < k1) c1 = vec_load(pidx, pval, k + 256, lane, SP);
        for (; k < k1; k += 256) {
          VecRegs c2;
          if (k + 512 < k1) c2 = vec_load(pidx, pval, k + 512, lane, SP);
          vec_consume(c0, k, k0, k1, row_first, row_last, lane, xb, yacc);
          c0 = c1;
          c1 = c2;
        }
      }
      __syncwarp();  // buffer b fully consumed (and tail stores of the staged buffer visible) before it is refilled
    }
    __syncthreads();
    const int32_t trow0 = P.tile_row0[t];
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < kWWarps; w++) acc += yacc_all[w * kTileRows + i];
      y[trow0 + i] = acc;
    }
    __syncthreads();
  }
}

__global__ void gather_perm_kernel(const double *x, const int32_t *idx, double *out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[idx[i]];
}
__global__ void add_perm_kernel(double *y, const double *yb, const int32_t *browL_inv, int64_t nloc) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nloc) y[i] = y[i] + yb[browL_inv[i]];
}

static unsigned long long *g_tile_counters = nullptr;  // [2]

int wcsr_spmv(sqmc_b200_handle *h, const double *x, double *y, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0;
  if (nloc == 0) return 0;
  if (!g_tile_counters) {
    SQ_CUDA(cudaMalloc(&g_tile_counters, 2 * sizeof(unsigned long long)));
    SQ_CUDA(cudaFuncSetAttribute(wcsr_spmv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWcsrSmemBytes));
  }
  SQ_CUDA(cudaMemsetAsync(g_tile_counters, 0, 2 * sizeof(unsigned long long), s));
  gather_perm_kernel<<<gblocks(h->n), 256, 0, s>>>(x, h->d_bidx, h->d_xb, h->n);
  SQ_LAUNCH_CHECK();
  const int grid = G.sm_count;
  wcsr_spmv_kernel<<<grid, kWThreads, kWcsrSmemBytes, s>>>(h->WA, h->d_cols, h->d_vals, x, h->n, y, g_tile_counters);
  SQ_LAUNCH_CHECK();
  wcsr_spmv_kernel<<<grid, kWThreads, kWcsrSmemBytes, s>>>(h->WB, h->d_cols, h->d_vals, h->d_xb, h->n, h->d_yb, g_tile_counters + 1);
  SQ_LAUNCH_CHECK();
  add_perm_kernel<<<gblocks(nloc), 256, 0, s>>>(y, h->d_yb, h->d_browL_inv, nloc);
  SQ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ conversion driver
static void free_part(WPart &P) {
  auto F = [](auto *&p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  F(P.tile_row0); F(P.tile_nrows); F(P.tile_ent0); F(P.tile_item0); F(P.items); F(P.rowptr); F(P.colwin);
  F(P.gwin_col0); F(P.gwin_len);
  P = WPart();
}
void wcsr_free(sqmc_b200_handle *h) {
  free_part(h->WA);
  free_part(h->WB);
  auto F = [](auto *&p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  F(h->d_browL); F(h->d_browL_inv); F(h->d_diag); F(h->d_xb); F(h->d_yb);
  h->wcsr = false;
}

template <typename T>
static int upload(T *&dev, const std::vector<T> &v) {
  SQ_CUDA(cudaMalloc(&dev, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) SQ_CUDA(cudaMemcpy(dev, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// global window table of one column space from its group offsets
static int make_window_table(WPart &P, const std::vector<int64_t> &goff, const int32_t *d_grp_of_col, const int64_t *d_goff, int64_t ncols, cudaStream_t s) {
  const int64_t ng = (int64_t)goff.size() - 1;
  std::vector<int32_t> wbase(ng + 1), c0, len;
  for (int64_t g = 0; g < ng; g++) {
    wbase[g] = (int32_t)c0.size();
    const int64_t sz = goff[g + 1] - goff[g];
    for (int64_t piece = 0; piece * kWinMax < sz; piece++) {
      int64_t a = goff[g] + piece * kWinMax, l = std::min<int64_t>(kWinMax, sz - piece * kWinMax);
      int64_t a0 = a & ~1ll, a1 = (a + l + 1) & ~1ll;
      c0.push_back((int32_t)a0);
      len.push_back((int32_t)(a1 - a0));
    }
  }
  wbase[ng] = (int32_t)c0.size();
  P.ngwin = (int64_t)c0.size();
  SQ_CHECK(upload(P.gwin_col0, c0));
  SQ_CHECK(upload(P.gwin_len, len));
  int32_t *d_wbase = nullptr;
  SQ_CHECK(upload(d_wbase, wbase));
  SQ_CUDA(cudaMalloc(&P.colwin, std::max<int64_t>(ncols, 1) * sizeof(int32_t)));
  colwin_kernel<<<gblocks(ncols), 256, 0, s>>>(d_grp_of_col, d_goff, d_wbase, P.colwin, ncols);
  SQ_LAUNCH_CHECK();
  SQ_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_wbase);
  return 0;
}

// tiles = runs of rows with equal group id, cut at kTileRows; grp[r] for part-local rows
static void make_tiles(const std::vector<int32_t> &grp, const std::vector<int64_t> &rowptr, std::vector<int32_t> &row0, std::vector<int32_t> &nrows,
                       std::vector<int64_t> &ent0) {
  const int64_t n = (int64_t)grp.size();
  int64_t r = 0;
  while (r < n) {
    int64_t e = r + 1;
    while (e < n && e - r < kTileRows && grp[e] == grp[r]) e++;
    row0.push_back((int32_t)r);
    nrows.push_back((int32_t)(e - r));
    ent0.push_back(rowptr[r]);
    r = e;
  }
  ent0.push_back(rowptr[n]);
}

static int convert_part(sqmc_b200_handle *h, WPart &P, const std::vector<int32_t> &grp_of_row, cudaStream_t s) {
  // rowptr (relative to P.ent0) is already on the device in P.rowptr
  std::vector<int64_t> rowptr(P.nrows + 1);
  SQ_CUDA(cudaMemcpy(rowptr.data(), P.rowptr, (P.nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  std::vector<int32_t> row0, nrows;
  std::vector<int64_t> ent0;
  make_tiles(grp_of_row, rowptr, row0, nrows, ent0);
  P.ntiles = (int64_t)row0.size();
  SQ_CHECK(upload(P.tile_row0, row0));
  SQ_CHECK(upload(P.tile_nrows, nrows));
  SQ_CHECK(upload(P.tile_ent0, ent0));
  int64_t max_tile_ent = 0;
  for (int64_t t = 0; t < P.ntiles; t++) max_tile_ent = std::max(max_tile_ent, ent0[t + 1] - ent0[t]);
  const int nwords = (int)div_up(P.ngwin, 32);
  // pass 1: work items per tile
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(P.ntiles, 1), G.sm_count * 2);
  DevBuf<int64_t> icount;
  DevBuf<int> maxw_dev;
  int max_w_tile = 0;
  SQ_CHECK(maxw_dev.alloc(1));
  SQ_CUDA(cudaMemsetAsync(maxw_dev.p, 0, sizeof(int), s));
  SQ_CHECK(icount.alloc(P.ntiles + 1));
  SQ_CUDA(cudaMemsetAsync(icount.p, 0, (P.ntiles + 1) * sizeof(int64_t), s));
  if (P.ntiles > 0) {
    DevBuf<int32_t> cntw;
    SQ_CHECK(cntw.alloc(std::max<int64_t>(P.ngwin, 1) * grid));
    SQ_CUDA(cudaMemsetAsync(cntw.p, 0, std::max<int64_t>(P.ngwin, 1) * grid * sizeof(int32_t), s));
    tile_count_items_kernel<<<grid, kWThreads, nwords * 4, s>>>(P, h->d_cols, nwords, cntw.p, icount.p, maxw_dev.p);
    SQ_LAUNCH_CHECK();
    SQ_CUDA(cudaMemcpyAsync(&max_w_tile, maxw_dev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SQ_CUDA(cudaStreamSynchronize(s));
  }
  std::vector<int64_t> ic(P.ntiles + 1, 0), item0(P.ntiles + 1, 0);
  SQ_CUDA(cudaMemcpy(ic.data(), icount.p, (P.ntiles + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  for (int64_t t = 0; t < P.ntiles; t++) item0[t + 1] = item0[t] + ic[t];
  P.nitems = item0[P.ntiles];
  SQ_CHECK(upload(P.tile_item0, item0));
  SQ_CUDA(cudaMalloc(&P.items, std::max<int64_t>(P.nitems, 1) * sizeof(WItem)));
  if (P.ntiles == 0) return 0;
  // pass 2: reorder.  The [windows][rows] count matrix lives in shared memory when it fits, else in global scratch
  // sized for the widest tile found in pass 1.
  int cnt_smem_ints = (int)((200 * 1024 - 2 * nwords * 4) / 4);
  if (cnt_smem_ints < 0) cnt_smem_ints = 0;
  int64_t max_w = std::max<int64_t>(max_w_tile, 1);
  const int64_t need = max_w * kTileRows;
  int smem_bytes = 2 * nwords * 4 + (int)std::min<int64_t>(need, cnt_smem_ints) * 4;
  if (need <= cnt_smem_ints) cnt_smem_ints = (int)need;
  DevBuf<int32_t> cnt_scr, scr_idx;
  DevBuf<double> scr_val;
  int64_t cnt_stride = (need > cnt_smem_ints) ? need : 1;
  SQ_CHECK(cnt_scr.alloc(cnt_stride * grid));
  SQ_CHECK(scr_idx.alloc(std::max<int64_t>(max_tile_ent, 1) * grid));
  SQ_CHECK(scr_val.alloc(std::max<int64_t>(max_tile_ent, 1) * grid));
  SQ_CUDA(cudaFuncSetAttribute(tile_reorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  tile_reorder_kernel<<<grid, kWThreads, smem_bytes, s>>>(P, h->d_cols, h->d_vals, nwords, cnt_smem_ints, cnt_scr.p, cnt_stride, scr_idx.p, scr_val.p,
                                                          std::max<int64_t>(max_tile_ent, 1));
  SQ_LAUNCH_CHECK();
  SQ_CUDA(cudaStreamSynchronize(s));
  // conversion-only tables are no longer needed
  cudaFree(P.rowptr); P.rowptr = nullptr;
  cudaFree(P.colwin); P.colwin = nullptr;
  cudaFree(P.gwin_col0); P.gwin_col0 = nullptr;
  cudaFree(P.gwin_len); P.gwin_len = nullptr;
  return 0;
}

template <int NW>
static int convert_impl(sqmc_b200_handle *h) {
  cudaStream_t s = G.stream;
  const int64_t n = h->n, nloc = h->row1 - h->row0, row0 = h->row0;
  // ---- diagonal (Davidson preconditioner) before the CSR disappears
  SQ_CUDA(cudaMalloc(&h->d_diag, std::max<int64_t>(nloc, 1) * sizeof(double)));
  if (nloc > 0) {
    extract_diag_kernel<<<gblocks(nloc), 256, 0, s>>>(h->d_rowptr, h->d_cols, h->d_vals, row0, nloc, h->d_diag);
    SQ_LAUNCH_CHECK();
  }
  // ---- local beta-major row order
  SQ_CUDA(cudaMalloc(&h->d_browL, std::max<int64_t>(nloc, 1) * sizeof(int32_t)));
  SQ_CUDA(cudaMalloc(&h->d_browL_inv, std::max<int64_t>(nloc, 1) * sizeof(int32_t)));
  {
    DevBuf<int32_t> num;
    SQ_CHECK(num.alloc(1));
    size_t tb = 0;
    cub::DeviceSelect::If(nullptr, tb, h->d_bidx, h->d_browL, num.p, (int)n, InRange{(int32_t)row0, (int32_t)h->row1}, s);
    DevBuf<char> tmp;
    SQ_CHECK(tmp.alloc((int64_t)tb + 16));
    SQ_CUDA(cub::DeviceSelect::If(tmp.p, tb, h->d_bidx, h->d_browL, num.p, (int)n, InRange{(int32_t)row0, (int32_t)h->row1}, s));
    g_launch_count += 2;
    if (nloc > 0) {
      invert_local_kernel<<<gblocks(nloc), 256, 0, s>>>(h->d_browL, h->d_browL_inv, row0, nloc);
      SQ_LAUNCH_CHECK();
    }
    SQ_CUDA(cudaStreamSynchronize(s));
  }
  // ---- classify entries, row pointers of the two parts
  DevBuf<int32_t> cntA, cntB;
  SQ_CHECK(cntA.alloc(nloc + 1));
  SQ_CHECK(cntB.alloc(nloc + 1));
  SQ_CUDA(cudaMemsetAsync(cntA.p, 0, (nloc + 1) * sizeof(int32_t), s));
  SQ_CUDA(cudaMemsetAsync(cntB.p, 0, (nloc + 1) * sizeof(int32_t), s));
  if (nloc > 0) {
    classify_kernel<NW><<<gblocks(nloc * 32), 256, 0, s>>>(h->d_rowptr, h->d_cols, h->d_dn, row0, nloc, cntA.p, cntB.p, h->d_browL_inv);
    SQ_LAUNCH_CHECK();
  }
  WPart &A = h->WA, &B = h->WB;
  A.nrows = B.nrows = nloc;
  SQ_CUDA(cudaMalloc(&A.rowptr, (nloc + 1) * sizeof(int64_t)));
  SQ_CUDA(cudaMalloc(&B.rowptr, (nloc + 1) * sizeof(int64_t)));
  {
    size_t tb = 0;
    auto itA = cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *>(cntA.p, cub::CastOp<int64_t>());
    auto itB = cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *>(cntB.p, cub::CastOp<int64_t>());
    cub::DeviceScan::ExclusiveSum(nullptr, tb, itA, A.rowptr, (int)(nloc + 1), s);
    DevBuf<char> tmp;
    SQ_CHECK(tmp.alloc((int64_t)tb + 16));
    SQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, itA, A.rowptr, (int)(nloc + 1), s));
    SQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, itB, B.rowptr, (int)(nloc + 1), s));
    g_launch_count += 4;
    SQ_CUDA(cudaStreamSynchronize(s));
  }
  std::vector<int64_t> rpA(nloc + 1), rp(nloc + 1);
  SQ_CUDA(cudaMemcpy(rpA.data(), A.rowptr, (nloc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  SQ_CUDA(cudaMemcpy(rp.data(), h->d_rowptr, (nloc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  int64_t nnzB = 0;
  SQ_CUDA(cudaMemcpy(&nnzB, B.rowptr + nloc, sizeof(int64_t), cudaMemcpyDeviceToHost));
  A.nnz = rpA[nloc];
  B.nnz = nnzB;
  A.ent0 = 0;
  B.ent0 = A.nnz;
  if (A.nnz + B.nnz != h->nnz_local) { set_error("wcsr: entry classification lost entries"); return 5; }
  // ---- split: B entries to a temporary, A entries compacted in place through a chunk scratch
  {
    DevBuf<int32_t> colsB, scrc;
    DevBuf<double> valsB, scrv;
    SQ_CHECK(colsB.alloc(std::max<int64_t>(nnzB, 1)));
    SQ_CHECK(valsB.alloc(std::max<int64_t>(nnzB, 1)));
    const int64_t kChunk = 1ll << 27;
    int64_t scr_cap = 0;
    int64_t r = 0;
    while (r < nloc) {
      int64_t limit = rp[r] + kChunk;
      int64_t r_end = std::upper_bound(rp.begin() + r + 1, rp.begin() + nloc + 1, limit) - rp.begin() - 1;
      if (r_end <= r) r_end = r + 1;
      if (r_end > nloc) r_end = nloc;
      int64_t na = rpA[r_end] - rpA[r];
      if (na > scr_cap) {
        scr_cap = na;
        SQ_CHECK(scrc.alloc(scr_cap));
        SQ_CHECK(scrv.alloc(scr_cap));
      }
      split_kernel<NW><<<gblocks((r_end - r) * 32), 256, 0, s>>>(h->d_rowptr, h->d_cols, h->d_vals, h->d_dn, row0, r, r_end, A.rowptr, B.rowptr,
                                                                h->d_browL_inv, h->d_binv, scrc.p, scrv.p, colsB.p, valsB.p);
      SQ_LAUNCH_CHECK();
      if (na > 0) {
        SQ_CUDA(cudaMemcpyAsync(h->d_cols + rpA[r], scrc.p, na * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        SQ_CUDA(cudaMemcpyAsync(h->d_vals + rpA[r], scrv.p, na * sizeof(double), cudaMemcpyDeviceToDevice, s));
      }
      SQ_CUDA(cudaStreamSynchronize(s));
      r = r_end;
    }
    if (nnzB > 0) {
      SQ_CUDA(cudaMemcpyAsync(h->d_cols + A.nnz, colsB.p, nnzB * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
      SQ_CUDA(cudaMemcpyAsync(h->d_vals + A.nnz, valsB.p, nnzB * sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    SQ_CUDA(cudaStreamSynchronize(s));
  }
  // ---- window tables and tiles
  std::vector<int64_t> gA(h->nA + 1), gB(h->nB + 1);
  SQ_CUDA(cudaMemcpy(gA.data(), h->d_gA_off, (h->nA + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  SQ_CUDA(cudaMemcpy(gB.data(), h->d_gB_off, (h->nB + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  SQ_CHECK(make_window_table(A, gA, h->d_eA, h->d_gA_off, n, s));
  SQ_CHECK(make_window_table(B, gB, h->d_eBpos, h->d_gB_off, n, s));
  std::vector<int32_t> grpA(nloc), grpB(nloc);
  if (nloc > 0) {
    SQ_CUDA(cudaMemcpy(grpA.data(), h->d_eA + row0, nloc * sizeof(int32_t), cudaMemcpyDeviceToHost));
    // beta group of local beta-major row i = eBpos[binv[browL[i]]]
    std::vector<int32_t> browL(nloc), binv(n), eBpos(n);
    SQ_CUDA(cudaMemcpy(browL.data(), h->d_browL, nloc * sizeof(int32_t), cudaMemcpyDeviceToHost));
    SQ_CUDA(cudaMemcpy(binv.data(), h->d_binv, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    SQ_CUDA(cudaMemcpy(eBpos.data(), h->d_eBpos, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < nloc; i++) grpB[i] = eBpos[binv[browL[i]]];
  }
  SQ_CHECK(convert_part(h, A, grpA, s));
  SQ_CHECK(convert_part(h, B, grpB, s));
  SQ_CUDA(cudaMalloc(&h->d_xb, std::max<int64_t>(n, 1) * sizeof(double)));
  SQ_CUDA(cudaMalloc(&h->d_yb, std::max<int64_t>(nloc, 1) * sizeof(double)));
  h->wcsr = true;
  return 0;
}

int wcsr_convert(sqmc_b200_handle *h) {
  // only for builds without time-reversal expansion (the group structure is kept then)
  if (!h->d_gA_off || !h->d_bidx || h->T.time_sym) return 0;
  // OPT-IN (SQMC_WCSR=1 forces the conversion, SQMC_WCSR=2 converts dense spaces only).  Default off: on B200 the
  // layout is correct and removes the L1TEX gather traffic, but its row-sum bookkeeping costs more issue slots
  // than the gathers it saves (profiles/r01_wcsr_experiment.txt: 36.7 ms vs 23.4 ms per H.v at 10^7 dets).
  const char *e = getenv("SQMC_WCSR");
  int mode = e ? atoi(e) : 0;
  if (mode == 2) mode = -1;
  if (mode == 0) return 0;
  int64_t ngw_bound = h->nA + h->n / kWinMax + 1, ngw_bound_b = h->nB + h->n / kWinMax + 1;
  if (ngw_bound > 65536 || ngw_bound_b > 65536) return 0;  // window bitmap of the conversion lives in shared memory
  if (mode < 0) {
    // dense spaces only: many determinants per alpha string and per beta string
    if (h->n < (1 << 16)) return 0;
    if (h->n / std::max<int64_t>(h->nA, 1) < 64 || h->n / std::max<int64_t>(h->nB, 1) < 64) return 0;
  }
  return h->NW == 1 ? convert_impl<1>(h) : convert_impl<2>(h);
}

// ------------------------------------------------------------------ single-row extraction (get_row on a WCSR matrix)
__global__ void wcsr_extract_row_kernel(WPart P, const int32_t *idx, const double *vals, int64_t t, int rl, int32_t *out_cols, double *out_vals,
                                        int *counter, int cap) {
  const int64_t e0 = P.tile_ent0[t];
  for (int64_t j = P.tile_item0[t]; j < P.tile_item0[t + 1]; j++) {
    const WItem it = P.items[j];
    for (int64_t k = e0 + it.k0 + threadIdx.x; k < e0 + it.k1; k += blockDim.x) {
      uint32_t u = (uint32_t)idx[P.ent0 + k];
      if ((int)(u >> 16) == rl) {
        int q = atomicAdd(counter, 1);
        if (q < cap) { out_cols[q] = it.col0 + (int32_t)(u & 0xffffu); out_vals[q] = vals[P.ent0 + k]; }
      }
    }
  }
}
// full row of internal row p (owned by this rank): internal column numbers, unsorted
int wcsr_get_row(sqmc_b200_handle *h, int64_t p, std::vector<int32_t> &cols, std::vector<double> &vals) {
  cudaStream_t s = G.stream;
  const int cap = 1 << 20;
  DevBuf<int32_t> oc;
  DevBuf<double> ov;
  DevBuf<int> cnt;
  SQ_CHECK(oc.alloc(cap));
  SQ_CHECK(ov.alloc(cap));
  SQ_CHECK(cnt.alloc(1));
  cols.clear();
  vals.clear();
  for (int part = 0; part < 2; part++) {
    WPart &P = part == 0 ? h->WA : h->WB;
    int64_t prow = p - h->row0;
    if (part == 1) {
      int32_t b = 0;
      SQ_CUDA(cudaMemcpy(&b, h->d_browL_inv + (p - h->row0), 4, cudaMemcpyDeviceToHost));
      prow = b;
    }
    if (P.ntiles == 0) continue;
    std::vector<int32_t> trow0(P.ntiles);
    SQ_CUDA(cudaMemcpy(trow0.data(), P.tile_row0, P.ntiles * 4, cudaMemcpyDeviceToHost));
    int64_t t = std::upper_bound(trow0.begin(), trow0.end(), (int32_t)prow) - trow0.begin() - 1;
    SQ_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int), s));
    wcsr_extract_row_kernel<<<1, 256, 0, s>>>(P, h->d_cols, h->d_vals, t, (int)(prow - trow0[t]), oc.p, ov.p, cnt.p, cap);
    SQ_LAUNCH_CHECK();
    int c = 0;
    SQ_CUDA(cudaMemcpyAsync(&c, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SQ_CUDA(cudaStreamSynchronize(s));
    if (c > cap) { set_error("wcsr_get_row: row longer than %d", cap); return 2; }
    std::vector<int32_t> hc(c);
    std::vector<double> hv(c);
    if (c) {
      SQ_CUDA(cudaMemcpy(hc.data(), oc.p, c * 4, cudaMemcpyDeviceToHost));
      SQ_CUDA(cudaMemcpy(hv.data(), ov.p, c * 8, cudaMemcpyDeviceToHost));
    }
    for (int k = 0; k < c; k++) {
      int32_t icol = hc[k];
      if (part == 1) SQ_CUDA(cudaMemcpy(&icol, h->d_bidx + hc[k], 4, cudaMemcpyDeviceToHost));
      cols.push_back(icol);
      vals.push_back(hv[k]);
    }
  }
  return 0;
}

// ------------------------------------------------------------------ host decode (export / get_row)
int wcsr_decode_host(sqmc_b200_handle *h, std::vector<int64_t> &rowptr, std::vector<int32_t> &cols, std::vector<double> &vals) {
  const int64_t nloc = h->row1 - h->row0, n = h->n;
  std::vector<int32_t> idx(h->nnz_local), bidx(n), browL(nloc);
  std::vector<double> v(h->nnz_local);
  SQ_CUDA(cudaMemcpy(idx.data(), h->d_cols, h->nnz_local * 4, cudaMemcpyDeviceToHost));
  SQ_CUDA(cudaMemcpy(v.data(), h->d_vals, h->nnz_local * 8, cudaMemcpyDeviceToHost));
  SQ_CUDA(cudaMemcpy(bidx.data(), h->d_bidx, n * 4, cudaMemcpyDeviceToHost));
  if (nloc) SQ_CUDA(cudaMemcpy(browL.data(), h->d_browL, nloc * 4, cudaMemcpyDeviceToHost));
  std::vector<std::vector<std::pair<int32_t, double>>> rows(nloc);
  for (int part = 0; part < 2; part++) {
    WPart &P = part == 0 ? h->WA : h->WB;
    if (P.ntiles == 0) continue;
    std::vector<int32_t> trow0(P.ntiles);
    std::vector<int64_t> tent0(P.ntiles + 1), titem0(P.ntiles + 1);
    std::vector<WItem> items(std::max<int64_t>(P.nitems, 1));
    SQ_CUDA(cudaMemcpy(trow0.data(), P.tile_row0, P.ntiles * 4, cudaMemcpyDeviceToHost));
    SQ_CUDA(cudaMemcpy(tent0.data(), P.tile_ent0, (P.ntiles + 1) * 8, cudaMemcpyDeviceToHost));
    SQ_CUDA(cudaMemcpy(titem0.data(), P.tile_item0, (P.ntiles + 1) * 8, cudaMemcpyDeviceToHost));
    if (P.nitems) SQ_CUDA(cudaMemcpy(items.data(), P.items, P.nitems * sizeof(WItem), cudaMemcpyDeviceToHost));
    for (int64_t t = 0; t < P.ntiles; t++) {
      for (int64_t j = titem0[t]; j < titem0[t + 1]; j++) {
        for (int64_t k = tent0[t] + items[j].k0; k < tent0[t] + items[j].k1; k++) {
          uint32_t u = (uint32_t)idx[P.ent0 + k];
          int64_t prow = trow0[t] + (u >> 16);
          int32_t pcol = items[j].col0 + (int32_t)(u & 0xffffu);
          int64_t lrow = part == 0 ? prow : (browL[prow] - h->row0);  // local alpha-major row
          int32_t icol = part == 0 ? pcol : bidx[pcol];               // internal column
          rows[lrow].push_back({icol, v[P.ent0 + k]});
        }
      }
    }
  }
  rowptr.assign(nloc + 1, 0);
  cols.clear();
  vals.clear();
  for (int64_t r = 0; r < nloc; r++) {
    std::sort(rows[r].begin(), rows[r].end(), [](const std::pair<int32_t, double> &a, const std::pair<int32_t, double> &b) { return a.first < b.first; });
    for (auto &e : rows[r]) { cols.push_back(e.first); vals.push_back(e.second); }
    rowptr[r + 1] = (int64_t)cols.size();
  }
  return 0;
}

}  // namespace sqmc
