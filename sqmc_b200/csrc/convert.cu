// convert.cu -- device-side conversion between the library's full symmetric CSR (internal row order) and the
// reference's layout (sparse_mat, commons/common_selected_ci.f90:26-31; what generate_sparse_ham_*_upper_triangular
// returns and what dtm_projector.* stores, do_walk.f90:883-1013): upper triangle only, rows in the CALLER's order,
// per-row counts, 1-based int64 columns ascending with the diagonal first.
//   export: filter (caller column >= caller row), renumber, per-row key/value sort in shared memory, widen to int64;
//           rows are processed in chunks so the temporaries stay bounded; results go straight to the caller's arrays.
//   import: upper -> full rows (each stored entry lands in its row and, if off-diagonal, in its column's row),
//           per-row key/value sort; internal order = caller order.
#include <cub/cub.cuh>

#include <algorithm>

#include "handle.h"

namespace sqmc {

static inline unsigned cblocks(int64_t n, int t = 256) { return (unsigned)std::max<int64_t>(1, div_up(n, t)); }

// ------------------------------------------------------------------ per-row key/value sort (all-ascending bitonic network)
__device__ __forceinline__ void bitonic_sort_kv(int32_t *k, double *v, int len, int tid, int nthr, bool block_sync) {
  int np2 = 1;
  while (np2 < len) np2 <<= 1;
  for (int sz = 2; sz <= np2; sz <<= 1) {
    for (int i = tid; i < np2; i += nthr) {
      int partner = i ^ (sz - 1);
      if (partner > i && partner < len && k[i] > k[partner]) {
        int32_t tk = k[i]; k[i] = k[partner]; k[partner] = tk;
        double tv = v[i]; v[i] = v[partner]; v[partner] = tv;
      }
    }
    if (block_sync) __syncthreads(); else __syncwarp();
    for (int j = sz >> 2; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += nthr) {
        int partner = i ^ j;
        if (partner > i && partner < len && k[i] > k[partner]) {
          int32_t tk = k[i]; k[i] = k[partner]; k[partner] = tk;
          double tv = v[i]; v[i] = v[partner]; v[partner] = tv;
        }
      }
      if (block_sync) __syncthreads(); else __syncwarp();
    }
  }
}
static const int kKvWarpMax = 128;     // rows up to this length: one warp, shared memory
static const int kKvBlockMax = 12288;  // rows up to this length: one CTA, shared memory (12 B per entry); longer: global memory

__global__ void __launch_bounds__(256) sort_kv_warp_kernel(const int64_t *ptr, int64_t nrows, int32_t *keys, double *vals) {
  __shared__ int32_t sk[8][kKvWarpMax];
  __shared__ double sv[8][kKvWarpMax];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = blockIdx.x * 8ll + w;
  if (row >= nrows) return;
  const int64_t b = ptr[row];
  const int L = (int)(ptr[row + 1] - b);
  if (L <= 1 || L > kKvWarpMax) return;
  for (int i = lane; i < L; i += 32) { sk[w][i] = keys[b + i]; sv[w][i] = vals[b + i]; }
  __syncwarp();
  bitonic_sort_kv(sk[w], sv[w], L, lane, 32, false);
  for (int i = lane; i < L; i += 32) { keys[b + i] = sk[w][i]; vals[b + i] = sv[w][i]; }
}
__global__ void __launch_bounds__(256) sort_kv_block_kernel(const int64_t *ptr, int64_t nrows, int32_t *keys, double *vals) {
  extern __shared__ __align__(8) unsigned char kvsm[];
  double *sv = reinterpret_cast<double *>(kvsm);
  int32_t *sk = reinterpret_cast<int32_t *>(sv + kKvBlockMax);
  for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int64_t b = ptr[row];
    const int L = (int)(ptr[row + 1] - b);
    if (L <= kKvWarpMax) continue;
    if (L <= kKvBlockMax) {
      for (int i = threadIdx.x; i < L; i += blockDim.x) { sk[i] = keys[b + i]; sv[i] = vals[b + i]; }
      __syncthreads();
      bitonic_sort_kv(sk, sv, L, threadIdx.x, blockDim.x, true);
      for (int i = threadIdx.x; i < L; i += blockDim.x) { keys[b + i] = sk[i]; vals[b + i] = sv[i]; }
      __syncthreads();
    } else {
      bitonic_sort_kv(keys + b, vals + b, L, threadIdx.x, blockDim.x, true);
    }
  }
}
static int sort_rows_kv(const int64_t *ptr, int64_t nrows, int64_t maxlen, int32_t *keys, double *vals, cudaStream_t s) {
  if (nrows == 0) return 0;
  sort_kv_warp_kernel<<<cblocks(nrows, 8), 256, 0, s>>>(ptr, nrows, keys, vals);
  SQ_LAUNCH_CHECK();
  if (maxlen > kKvWarpMax) {
    const int smem = kKvBlockMax * 12;
    SQ_CUDA(cudaFuncSetAttribute(sort_kv_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    sort_kv_block_kernel<<<(unsigned)std::min<int64_t>(nrows, G.sm_count * 8), 256, smem, s>>>(ptr, nrows, keys, vals);
    SQ_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------ export
__global__ void __launch_bounds__(256) export_count_kernel(const int64_t *rowptr, const int32_t *cols, const int32_t *perm, int64_t row0, int64_t nloc,
                                                           int32_t *cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= nloc) return;
  const int32_t ci = perm[row0 + r];
  int c = 0;
  for (int64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) c += perm[cols[k]] >= ci;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cnt[r] = c;
}
__global__ void gather_counts_kernel(const int32_t *cnt, const int32_t *order, int64_t n, int64_t *out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = cnt[order[i]];
}
// output rows [k0,k1) (ascending caller index): write (caller column, value) of the kept entries
__global__ void __launch_bounds__(256) export_fill_kernel(const int64_t *rowptr, const int32_t *cols, const double *vals, const int32_t *perm, int64_t row0,
                                                          const int32_t *order, int64_t k0, int64_t k1, const int64_t *optr, int64_t obase,
                                                          int32_t *ocols, double *ovals) {
  const int lane = threadIdx.x & 31;
  const int64_t kk = k0 + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (kk >= k1) return;
  const int32_t r = order[kk];
  const int32_t ci = perm[row0 + r];
  int64_t w = optr[kk] - obase;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t b = rowptr[r], e = rowptr[r + 1];
  for (int64_t kb = b; kb < e; kb += 32) {
    const int64_t k = kb + lane;
    const bool in = k < e;
    const int32_t cj = in ? perm[cols[k]] : -1;
    const bool keep = in && cj >= ci;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int64_t q = w + __popc(m & lt);
      ocols[q] = cj;
      ovals[q] = vals[k];
    }
    w += __popc(m);
  }
}
__global__ void widen_kernel(const int32_t *in, int64_t *out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)in[i] + 1;
}
__global__ void iota32_kernel(int32_t *a, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)i;
}
__global__ void gather_perm32_kernel(const int32_t *perm, int64_t row0, int64_t n, int32_t *out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = perm[row0 + i];
}

int export_upper_device(sqmc_b200_handle *h, int64_t *counts, int64_t *indices, double *values) {
  cudaStream_t s = G.stream;
  const int64_t nloc = h->row1 - h->row0;
  if (nloc == 0) return 0;
  // output row order: local rows by ascending caller index
  DevBuf<int32_t> key_in, key_out, ord_in, order, cnt;
  SQ_CHECK(key_in.alloc(nloc)); SQ_CHECK(key_out.alloc(nloc)); SQ_CHECK(ord_in.alloc(nloc)); SQ_CHECK(order.alloc(nloc)); SQ_CHECK(cnt.alloc(nloc));
  gather_perm32_kernel<<<cblocks(nloc), 256, 0, s>>>(h->d_perm, h->row0, nloc, key_in.p);
  SQ_LAUNCH_CHECK();
  iota32_kernel<<<cblocks(nloc), 256, 0, s>>>(ord_in.p, nloc);
  SQ_LAUNCH_CHECK();
  {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key_in.p, key_out.p, ord_in.p, order.p, (int)nloc, 0, 32, s);
    DevBuf<char> tmp;
    SQ_CHECK(tmp.alloc((int64_t)tb + 16));
    SQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key_in.p, key_out.p, ord_in.p, order.p, (int)nloc, 0, 32, s));
    g_launch_count += 3;
    SQ_CUDA(cudaStreamSynchronize(s));
  }
  export_count_kernel<<<cblocks(nloc * 32), 256, 0, s>>>(h->d_rowptr, h->d_cols, h->d_perm, h->row0, nloc, cnt.p);
  SQ_LAUNCH_CHECK();
  DevBuf<int64_t> ocnt, optr;
  SQ_CHECK(ocnt.alloc(nloc + 1));
  SQ_CHECK(optr.alloc(nloc + 1));
  SQ_CUDA(cudaMemsetAsync(ocnt.p, 0, (nloc + 1) * sizeof(int64_t), s));
  gather_counts_kernel<<<cblocks(nloc), 256, 0, s>>>(cnt.p, order.p, nloc, ocnt.p);
  SQ_LAUNCH_CHECK();
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, ocnt.p, optr.p, (int)(nloc + 1), s);
    DevBuf<char> tmp;
    SQ_CHECK(tmp.alloc((int64_t)tb + 16));
    SQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, ocnt.p, optr.p, (int)(nloc + 1), s));
    g_launch_count += 2;
  }
  std::vector<int64_t> hptr(nloc + 1);
  SQ_CUDA(cudaMemcpyAsync(hptr.data(), optr.p, (nloc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  SQ_CUDA(cudaMemcpyAsync(counts, ocnt.p, nloc * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  SQ_CUDA(cudaStreamSynchronize(s));
  // chunks of output rows
  const int64_t kChunk = 1ll << 26;
  DevBuf<int32_t> tc;
  DevBuf<double> tv;
  DevBuf<int64_t> ti;
  int64_t cap = 0;
  int64_t k0 = 0;
  while (k0 < nloc) {
    int64_t k1 = std::upper_bound(hptr.begin() + k0 + 1, hptr.end(), hptr[k0] + kChunk) - hptr.begin() - 1;
    if (k1 <= k0) k1 = k0 + 1;
    if (k1 > nloc) k1 = nloc;
    const int64_t m = hptr[k1] - hptr[k0];
    if (m > cap) {
      cap = m;
      SQ_CHECK(tc.alloc(cap)); SQ_CHECK(tv.alloc(cap)); SQ_CHECK(ti.alloc(cap));
    }
    int64_t maxlen = 0;
    for (int64_t k = k0; k < k1; k++) maxlen = std::max(maxlen, hptr[k + 1] - hptr[k]);
    export_fill_kernel<<<cblocks((k1 - k0) * 32), 256, 0, s>>>(h->d_rowptr, h->d_cols, h->d_vals, h->d_perm, h->row0, order.p, k0, k1, optr.p, hptr[k0],
                                                              tc.p, tv.p);
    SQ_LAUNCH_CHECK();
    // per-row sort by caller column: row pointers relative to the chunk
    DevBuf<int64_t> cptr;
    SQ_CHECK(cptr.alloc(k1 - k0 + 1));
    std::vector<int64_t> rel(hptr.begin() + k0, hptr.begin() + k1 + 1);
    for (auto &v : rel) v -= hptr[k0];
    SQ_CUDA(cudaMemcpyAsync(cptr.p, rel.data(), rel.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    SQ_CHECK(sort_rows_kv(cptr.p, k1 - k0, maxlen, tc.p, tv.p, s));
    if (m > 0) {
      widen_kernel<<<cblocks(m), 256, 0, s>>>(tc.p, ti.p, m);
      SQ_LAUNCH_CHECK();
      SQ_CUDA(cudaMemcpyAsync(indices + hptr[k0], ti.p, m * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
      SQ_CUDA(cudaMemcpyAsync(values + hptr[k0], tv.p, m * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    SQ_CUDA(cudaStreamSynchronize(s));
    k0 = k1;
  }
  return 0;
}

// ------------------------------------------------------------------ import
__global__ void expand_rows_kernel(const int64_t *uptr, int64_t n, int32_t *row_of) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  for (int64_t k = uptr[r] + lane; k < uptr[r + 1]; k += 32) row_of[k] = (int32_t)r;
}
__global__ void import_degree_kernel(const int64_t *uidx, const int32_t *row_of, int64_t nnzu, int64_t n, int32_t *deg, int *bad) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= nnzu) return;
  const int64_t c = uidx[k] - 1;
  if (c < 0 || c >= n) { atomicExch(bad, 1); return; }
  const int32_t i = row_of[k];
  atomicAdd(&deg[i], 1);
  if (c != i) atomicAdd(&deg[c], 1);
}
__global__ void import_fill_kernel(const int64_t *uidx, const double *uval, const int32_t *row_of, int64_t nnzu, const int64_t *rowptr, int32_t *cursor,
                                   int32_t *cols, double *vals) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= nnzu) return;
  const int32_t c = (int32_t)(uidx[k] - 1), i = row_of[k];
  const double v = uval[k];
  int64_t q = rowptr[i] + atomicAdd(&cursor[i], 1);
  cols[q] = c;
  vals[q] = v;
  if (c != i) {
    q = rowptr[c] + atomicAdd(&cursor[c], 1);
    cols[q] = i;
    vals[q] = v;
  }
}

int import_upper_device(sqmc_b200_handle *h, int64_t n, const int64_t *counts, const int64_t *indices, const double *values) {
  cudaStream_t s = G.stream;
  int64_t nnzu = 0, maxu = 0;
  for (int64_t i = 0; i < n; i++) {
    if (counts[i] < 0) { set_error("import_upper: negative row count"); return 2; }
    nnzu += counts[i];
    maxu = std::max(maxu, counts[i]);
  }
  DevBuf<int64_t> ucnt, uptr, uidx;
  DevBuf<double> uval;
  DevBuf<int32_t> row_of, deg, cursor;
  DevBuf<int> bad;
  SQ_CHECK(ucnt.alloc(n + 1)); SQ_CHECK(uptr.alloc(n + 1)); SQ_CHECK(uidx.alloc(std::max<int64_t>(nnzu, 1))); SQ_CHECK(uval.alloc(std::max<int64_t>(nnzu, 1)));
  SQ_CHECK(row_of.alloc(std::max<int64_t>(nnzu, 1))); SQ_CHECK(deg.alloc(n + 1)); SQ_CHECK(cursor.alloc(n + 1)); SQ_CHECK(bad.alloc(1));
  SQ_CUDA(cudaMemsetAsync(ucnt.p, 0, (n + 1) * sizeof(int64_t), s));
  SQ_CUDA(cudaMemcpyAsync(ucnt.p, counts, n * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  SQ_CUDA(cudaMemcpyAsync(uidx.p, indices, nnzu * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  SQ_CUDA(cudaMemcpyAsync(uval.p, values, nnzu * sizeof(double), cudaMemcpyHostToDevice, s));
  SQ_CUDA(cudaMemsetAsync(deg.p, 0, (n + 1) * sizeof(int32_t), s));
  SQ_CUDA(cudaMemsetAsync(cursor.p, 0, (n + 1) * sizeof(int32_t), s));
  SQ_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, ucnt.p, uptr.p, (int)(n + 1), s);
    DevBuf<char> tmp;
    SQ_CHECK(tmp.alloc((int64_t)tb + 16));
    SQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, ucnt.p, uptr.p, (int)(n + 1), s));
    g_launch_count += 2;
  }
  expand_rows_kernel<<<cblocks(n * 32), 256, 0, s>>>(uptr.p, n, row_of.p);
  SQ_LAUNCH_CHECK();
  if (nnzu > 0) {
    import_degree_kernel<<<cblocks(nnzu), 256, 0, s>>>(uidx.p, row_of.p, nnzu, n, deg.p, bad.p);
    SQ_LAUNCH_CHECK();
  }
  int hbad = 0;
  SQ_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  SQ_CUDA(cudaStreamSynchronize(s));
  if (hbad) { set_error("import_upper: column out of range"); return 2; }
  SQ_CHECK(devbuf_alloc((void **)&h->d_rowptr, (n + 1) * sizeof(int64_t)));
  {
    size_t tb = 0;
    auto it = cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *>(deg.p, cub::CastOp<int64_t>());
    cub::DeviceScan::ExclusiveSum(nullptr, tb, it, h->d_rowptr, (int)(n + 1), s);
    DevBuf<char> tmp;
    SQ_CHECK(tmp.alloc((int64_t)tb + 16));
    SQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, h->d_rowptr, (int)(n + 1), s));
    g_launch_count += 2;
  }
  int64_t nnzf = 0;
  SQ_CUDA(cudaMemcpyAsync(&nnzf, h->d_rowptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  std::vector<int32_t> hdeg(n);
  SQ_CUDA(cudaMemcpyAsync(hdeg.data(), deg.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  SQ_CUDA(cudaStreamSynchronize(s));
  int64_t maxlen = 0;
  for (int64_t i = 0; i < n; i++) maxlen = std::max<int64_t>(maxlen, hdeg[i]);
  h->capacity = std::max<int64_t>(nnzf, 1);
  SQ_CHECK(matrix_arrays_ensure(h, h->capacity));
  if (nnzu > 0) {
    import_fill_kernel<<<cblocks(nnzu), 256, 0, s>>>(uidx.p, uval.p, row_of.p, nnzu, h->d_rowptr, cursor.p, h->d_cols, h->d_vals);
    SQ_LAUNCH_CHECK();
  }
  SQ_CHECK(sort_rows_kv(h->d_rowptr, n, maxlen, h->d_cols, h->d_vals, s));
  SQ_CUDA(cudaStreamSynchronize(s));
  h->nnz_local = nnzf;
  h->nnz_full = nnzf;
  h->nnz_upper = nnzu;
  return 0;
}

}  // namespace sqmc
