// spmv.cu -- y = H x on the device-resident full symmetric CSR (sm_100a).
//
// Replaces fast_sparse_matrix_multiply_upper_triangular (more_tools.f90:3622-3670),
// its MPI twin (:3674-3725) and the column-band variant (:3562-3587).  The reference
// walks the upper triangle with a running index and scatters answer(m) += H*v(i)
// (a serial loop dependence); here every row of the full matrix is owned by one
// sub-warp / warp / CTA, so there is no scatter and no atomics, and the summation
// order of a row is fixed (deterministic results).
//
// HBM-bound: 12 B per stored entry (f64 value + i32 column) streamed once with
// no-L1-allocate / L2 evict-first loads so that x (8 B per row, gathered) stays resident in the 126 MB L2.
// Rows are binned by degree: sub-warp vectors of 2/4/8/16/32 lanes per row, and one
// CTA per row for very long rows.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "handle.h"
#include "stream_loads.cuh"

namespace sqmc {

__device__ __forceinline__ int bin_of_len(int64_t len) {
  if (len <= 8) return 0;
  if (len <= 16) return 1;
  if (len <= 32) return 2;
  if (len <= 64) return 3;
  if (len <= 4096) return 4;
  return 5;
}
struct BinPred {
  const int64_t *rowptr;
  int bin;
  __device__ __forceinline__ bool operator()(const int32_t &row) const { return bin_of_len(rowptr[row + 1] - rowptr[row]) == bin; }
};

int spmv_setup_bins(sqmc_b200_handle *h) {
  cudaStream_t s = G.stream;
  const int64_t nloc = h->row1 - h->row0;
  if (h->d_bin_rows) devbuf_free(h->d_bin_rows);
  h->d_bin_rows = nullptr;
  for (int b = 0; b <= kNumBins; b++) h->bin_off[b] = 0;
  h->bins_ready = nloc == 0;
  if (nloc == 0) return 0;
  SQ_CHECK(devbuf_alloc((void **)&h->d_bin_rows, nloc * sizeof(int32_t)));
  int32_t *d_num = nullptr;
  SQ_CUDA(cudaMalloc(&d_num, sizeof(int32_t)));
  cub::CountingInputIterator<int32_t> it(0);
  size_t tb = 0;
  cub::DeviceSelect::If(nullptr, tb, it, h->d_bin_rows, d_num, (int)nloc, BinPred{h->d_rowptr, 0}, s);
  void *tmp = nullptr;
  SQ_CUDA(cudaMalloc(&tmp, tb + 16));
  int64_t off = 0;
  for (int b = 0; b < kNumBins; b++) {
    h->bin_off[b] = off;
    SQ_CUDA(cub::DeviceSelect::If(tmp, tb, it, h->d_bin_rows + off, d_num, (int)nloc, BinPred{h->d_rowptr, b}, s));
    g_launch_count += 2;
    int32_t num = 0;
    SQ_CUDA(cudaMemcpyAsync(&num, d_num, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    SQ_CUDA(cudaStreamSynchronize(s));
    off += num;
  }
  h->bin_off[kNumBins] = off;
  cudaFree(tmp);
  cudaFree(d_num);
  if (off != nloc) {
    set_error("spmv_setup_bins: bins cover %lld of %lld rows", (long long)off, (long long)nloc);
    return 4;
  }
  h->bins_ready = true;
  return 0;
}

template <int VS>
__global__ void __launch_bounds__(256) spmv_vec_kernel(const int32_t *__restrict__ rows, int64_t nrows, const int64_t *__restrict__ rowptr,
                                                       const int32_t *__restrict__ cols, const double *__restrict__ vals,
                                                       const double *__restrict__ x, double *__restrict__ y) {
  const Policies P;
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / VS;
  const int l = threadIdx.x & (VS - 1);
  const bool valid = gid < nrows;
  int32_t row = 0;
  int64_t k = 0, e = 0;
  if (valid) {
    row = rows[gid];
    k = rowptr[row] + l;
    e = rowptr[row + 1];
  }
  double acc = 0.0;
  // 4-way unrolled: the four column loads and four value loads are independent, then four gathers
  for (; k + 3 * VS < e; k += 4 * VS) {
    int32_t c0 = ld_col(cols + k, P), c1 = ld_col(cols + k + VS, P), c2 = ld_col(cols + k + 2 * VS, P), c3 = ld_col(cols + k + 3 * VS, P);
    double v0 = ld_val(vals + k, P), v1 = ld_val(vals + k + VS, P), v2 = ld_val(vals + k + 2 * VS, P), v3 = ld_val(vals + k + 3 * VS, P);
    double x0 = ld_x(x + c0, P), x1 = ld_x(x + c1, P), x2 = ld_x(x + c2, P), x3 = ld_x(x + c3, P);
    acc += v0 * x0;
    acc += v1 * x1;
    acc += v2 * x2;
    acc += v3 * x3;
  }
  for (; k < e; k += VS) acc += ld_val(vals + k, P) * ld_x(x + ld_col(cols + k, P), P);
#pragma unroll
  for (int o = VS >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, VS);
  if (valid && l == 0) y[row] = acc;
}

__global__ void __launch_bounds__(256) spmv_cta_kernel(const int32_t *__restrict__ rows, int64_t nrows, const int64_t *__restrict__ rowptr,
                                                       const int32_t *__restrict__ cols, const double *__restrict__ vals,
                                                       const double *__restrict__ x, double *__restrict__ y) {
  __shared__ double part[8];
  const Policies P;
  for (int64_t gid = blockIdx.x; gid < nrows; gid += gridDim.x) {
    const int32_t row = rows[gid];
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    double acc = 0.0;
    for (int64_t k = s + threadIdx.x; k < e; k += blockDim.x) acc += ld_val(vals + k, P) * ld_x(x + ld_col(cols + k, P), P);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
      y[row] = t;
    }
    __syncthreads();
  }
}

template <int VS>
static int launch_vec(sqmc_b200_handle *h, int bin, const double *x, double *y, cudaStream_t s) {
  int64_t nr = h->bin_off[bin + 1] - h->bin_off[bin];
  if (nr == 0) return 0;
  int64_t threads = nr * VS;
  spmv_vec_kernel<VS><<<(unsigned)div_up(threads, 256), 256, 0, s>>>(h->d_bin_rows + h->bin_off[bin], nr, h->d_rowptr, h->d_cols, h->d_vals, x, y);
  SQ_LAUNCH_CHECK();
  return 0;
}
static int launch_vec_bins(sqmc_b200_handle *h, const double *x, double *y, cudaStream_t s) {
  SQ_CHECK((launch_vec<2>(h, 0, x, y, s)));
  SQ_CHECK((launch_vec<4>(h, 1, x, y, s)));
  SQ_CHECK((launch_vec<8>(h, 2, x, y, s)));
  SQ_CHECK((launch_vec<16>(h, 3, x, y, s)));
  SQ_CHECK((launch_vec<32>(h, 4, x, y, s)));
  return 0;
}
static int launch_cta_bin(sqmc_b200_handle *h, const double *x, double *y, cudaStream_t s) {
  int64_t nr = h->bin_off[6] - h->bin_off[5];
  if (nr > 0) {
    spmv_cta_kernel<<<(unsigned)std::min<int64_t>(nr, 148 * 8), 256, 0, s>>>(h->d_bin_rows + h->bin_off[5], nr, h->d_rowptr, h->d_cols, h->d_vals, x, y);
    SQ_LAUNCH_CHECK();
  }
  return 0;
}

// y_dev: local rows (row1-row0), x_dev: global length n, both internal order
int spmv_launch(sqmc_b200_handle *h, const double *x, double *y, cudaStream_t s) {
  if (!h->d_rowptr) { set_error("matvec: no matrix on this handle"); return 2; }
  if (h->bundle_R) return bundle_spmv(h, x, y, s);
  if (!h->bins_ready) SQ_CHECK(spmv_setup_bins(h));  // degree bins of the plain-row kernels: built on first use
  SQ_CHECK(launch_vec_bins(h, x, y, s));
  return launch_cta_bin(h, x, y, s);
}

// Make the whole vector (k interleaved vectors) available on this rank.  `block` = this rank's row block
// (nloc*k doubles).  One GPU: nothing to do.  Several GPUs: every rank stores its block into every rank's x buffer
// over NVLink (csrc/p2p.cu), or -- when peer memory could not be mapped -- NCCL all-gathers into the handle's buffer.
int gather_vector(sqmc_b200_handle *h, const double *block, int k, cudaStream_t s, const double **full) {
  const int64_t nloc = h->row1 - h->row0;
  if (G.nranks == 1) { *full = block; return 0; }
  if (h->p2p.on) {
    double *f = nullptr;
    SQ_CHECK(p2p_gather(h, block, nloc * k, h->row0 * k, 0, s, &f));
    *full = f;
    return 0;
  }
  double *buf = h->d_x;
  if (k == 2) {
    if (!h->d_x2) SQ_CHECK(devbuf_alloc((void **)&h->d_x2, std::max<int64_t>(h->n, 1) * 2 * sizeof(double)));
    buf = h->d_x2;
  }
  if (nloc > 0 && block != buf + h->row0 * k) SQ_CUDA(cudaMemcpyAsync(buf + h->row0 * k, block, nloc * k * sizeof(double), cudaMemcpyDeviceToDevice, s));
  SQ_CHECK(allgather_rows_k(h, buf, k, s));
  *full = buf;
  return 0;
}

// y = H x with x given as this rank's row block
int spmv_block(sqmc_b200_handle *h, const double *block, double *y, cudaStream_t s) {
  if (!h->d_rowptr) { set_error("matvec: no matrix on this handle"); return 2; }
  const double *full = nullptr;
  SQ_CHECK(gather_vector(h, block, 1, s, &full));
  return spmv_launch(h, full, y, s);
}

// ---------------------------------------------------------------- diagonal of the local rows
__global__ void extract_diag_kernel(const int64_t *rowptr, const int32_t *cols, const double *vals, int64_t row0, int64_t nloc, double *diag) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= nloc) return;
  int64_t lo = rowptr[q], hi = rowptr[q + 1];
  const int32_t target = (int32_t)(row0 + q);
  double d = 0.0;
  while (lo < hi) {  // columns of a plain row are ascending
    int64_t mid = (lo + hi) >> 1;
    int32_t c = cols[mid];
    if (c == target) { d = vals[mid]; break; }
    if (c < target) lo = mid + 1;
    else hi = mid;
  }
  diag[q] = d;
}
// the copy kept in the handle when the entries are re-ordered (bundle.cu), else read from the plain rows
int extract_diag(sqmc_b200_handle *h, double *diag_dev, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0;
  if (nloc == 0) return 0;
  if (h->d_diag) {
    SQ_CUDA(cudaMemcpyAsync(diag_dev, h->d_diag, nloc * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  if (h->bundle_R) { set_error("extract_diag: re-ordered matrix without a kept diagonal"); return 4; }
  extract_diag_kernel<<<(unsigned)div_up(nloc, 256), 256, 0, s>>>(h->d_rowptr, h->d_cols, h->d_vals, h->row0, nloc, diag_dev);
  SQ_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- small vector helpers
__global__ void gather_kernel(const double *src, const int32_t *idx, double *dst, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void scatter_kernel(const double *src, const int32_t *idx, double *dst, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[i];
}
__global__ void scale_kernel(double *a, int64_t n, double r) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = a[i] * r;
}
__global__ void axpy_kernel(double *y, const double *x, double c, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = y[i] + c * x[i];
}
__global__ void gather_i32_kernel(const int32_t *src, const int32_t *idx, int32_t *dst, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
int gather_i32(const int32_t *src, const int32_t *idx, int32_t *dst, int64_t n, cudaStream_t s) {
  if (n == 0) return 0;
  gather_i32_kernel<<<(unsigned)div_up(n, 256), 256, 0, s>>>(src, idx, dst, n);
  SQ_LAUNCH_CHECK();
  return 0;
}
int permute_gather(const double *src, const int32_t *idx, double *dst, int64_t n, cudaStream_t s) {
  if (n == 0) return 0;
  gather_kernel<<<(unsigned)div_up(n, 256), 256, 0, s>>>(src, idx, dst, n);
  SQ_LAUNCH_CHECK();
  return 0;
}
int permute_scatter(const double *src, const int32_t *idx, double *dst, int64_t n, cudaStream_t s) {
  if (n == 0) return 0;
  scatter_kernel<<<(unsigned)div_up(n, 256), 256, 0, s>>>(src, idx, dst, n);
  SQ_LAUNCH_CHECK();
  return 0;
}
int scale_array(double *a, int64_t n, double r, cudaStream_t s) {
  if (n == 0) return 0;
  scale_kernel<<<(unsigned)div_up(n, 256), 256, 0, s>>>(a, n, r);
  SQ_LAUNCH_CHECK();
  return 0;
}
int projector_epilogue(double *deltaw, const double *w, double c, int64_t n, cudaStream_t s) {
  if (n == 0) return 0;
  axpy_kernel<<<(unsigned)div_up(n, 256), 256, 0, s>>>(deltaw, w, c, n);
  SQ_LAUNCH_CHECK();
  return 0;
}

// in-place all-gather of the rank-owned row blocks of a global-length vector
// (the reference emulates this with a zero-padded MPI_ALLREDUCE, more_tools.f90:2647,2772)
int allgather_rows(sqmc_b200_handle *h, double *x_full, cudaStream_t s) { return allgather_rows_k(h, x_full, 1, s); }

int allgather_rows_k(sqmc_b200_handle *h, double *x_full, int k, cudaStream_t s) {
  if (G.nranks == 1) return 0;
  ncclGroupStart();
  for (int r = 0; r < G.nranks; r++) {
    int64_t c = (h->row_starts[r + 1] - h->row_starts[r]) * k;
    if (c == 0) continue;
    ncclBroadcast(x_full + h->row_starts[r] * k, x_full + h->row_starts[r] * k, c, ncclDouble, r, G.comm, s);
  }
  ncclResult_t r = ncclGroupEnd();
  if (r != ncclSuccess) { set_error("allgather_rows_k: NCCL error %s", ncclGetErrorString(r)); return 3; }
  return 0;
}

__global__ void interleave2_kernel(const double *a, const double *b, double *out2, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { out2[2 * i] = a[i]; out2[2 * i + 1] = b[i]; }
}
__global__ void deinterleave2_kernel(const double *in2, double *a, double *b, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { a[i] = in2[2 * i]; b[i] = in2[2 * i + 1]; }
}

int spmv_pair(sqmc_b200_handle *h, const double *Va, const double *Vb, double *HVa, double *HVb, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0;
  if (!h->bundle_R) {  // plain rows: two single-vector products
    SQ_CHECK(spmv_block(h, Va, HVa, s));
    return spmv_block(h, Vb, HVb, s);
  }
  if (!h->d_y2) {
    SQ_CHECK(devbuf_alloc((void **)&h->d_xi2, std::max<int64_t>(nloc, 1) * 2 * sizeof(double)));
    SQ_CHECK(devbuf_alloc((void **)&h->d_y2, std::max<int64_t>(nloc, 1) * 2 * sizeof(double)));
  }
  if (nloc > 0) {
    interleave2_kernel<<<(unsigned)div_up(nloc, 256), 256, 0, s>>>(Va, Vb, h->d_xi2, nloc);
    SQ_LAUNCH_CHECK();
  }
  const double *full = nullptr;
  SQ_CHECK(gather_vector(h, h->d_xi2, 2, s, &full));
  SQ_CHECK(bundle_spmm2(h, full, h->d_y2, s));
  if (nloc > 0) {
    deinterleave2_kernel<<<(unsigned)div_up(nloc, 256), 256, 0, s>>>(h->d_y2, HVa, HVb, nloc);
    SQ_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace sqmc
