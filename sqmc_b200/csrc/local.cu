// local.cu -- the caller's data distribution: every MPI rank owns the determinants a hash assigns to it
// (get_det_owner, mpi_routines.f90:419-445) and passes / receives only its own slice of a vector, in ascending caller
// index: walk_wt(my_locations_of_imp_dets(1:my_nimp)) in, the reduce-scattered deltaw(1:my_nimp) out
// (do_walk.f90:2259-2260, mpi_routines.f90:1592); davidson_sparse_mpi2 works on local_det_map%ndets-long vectors
// (more_tools.f90:2525, 2842-2861).
//
// The library keeps its own row sharding (contiguous blocks of the alpha-major order); this file moves vectors between
// the two distributions on the device:
//   slices in  : H2D of the owned entries only -> every rank stores its slice into all ranks' staging buffer over NVLink
//                (slice-concatenated order: rank 0's determinants, then rank 1's, ... -- the order the reference's
//                band shuffle produces, more_tools.f90:3371-3471) -> one local permutation into internal order;
//   slices out : every rank scatters its result rows straight into the owners' buffers (p2p.cu) -> D2H of the owned
//                entries only.
// With one rank the same entry points work on full vectors.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "handle.h"

namespace sqmc {

static inline unsigned lblocks(int64_t n) { return (unsigned)std::max<int64_t>(1, div_up(n, 256)); }

__global__ void liota_kernel(int32_t *a, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)i;
}
// shuf_pos[sorted_row[j]] = j
__global__ void invert_kernel(const int32_t *sorted_row, int32_t *shuf_pos, int64_t n) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < n) shuf_pos[sorted_row[j]] = (int32_t)j;
}
__global__ void shuf_of_internal_kernel(const int32_t *perm, const int32_t *shuf_pos, int32_t *out, int64_t n) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < n) out[p] = shuf_pos[perm[p]];
}
struct OffTable {
  int64_t off[kMaxRanks];
};
__global__ void dest_kernel(const int32_t *perm, const int32_t *owner, const int32_t *shuf_pos, OffTable T, int64_t row0, int64_t nloc, int32_t *dest_rank,
                            int32_t *dest_pos) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= nloc) return;
  const int32_t i = perm[row0 + q];
  const int32_t r = owner[i];
  dest_rank[q] = r;
  dest_pos[q] = (int32_t)(shuf_pos[i] - T.off[r]);
}
__global__ void my_internal_kernel(const int32_t *sorted_row, const int32_t *iperm, int64_t off, int64_t my_n, int32_t *out) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < my_n) out[k] = iperm[sorted_row[off + k]];
}

int set_ownership(sqmc_b200_handle *h, const int32_t *owner_host, int64_t *n_owned_out) {
  if (!h->d_perm || h->n <= 0) { set_error("set_ownership: build or import the matrix first (the map is per determinant list)"); return 2; }
  cudaStream_t s = G.stream;
  const int64_t n = h->n, nloc = h->row1 - h->row0;
  const int R = G.nranks;
  h->own_count.assign(R, 0);
  h->own_off.assign(R + 1, 0);
  for (int64_t i = 0; i < n; i++) {
    const int32_t r = owner_host[i];
    if (r < 0 || r >= R) { set_error("set_ownership: owner_of_row(%lld) = %d is not a rank in 0..%d", (long long)i + 1, r, R - 1); h->own_set = false; return 2; }
    h->own_count[r]++;
  }
  for (int r = 0; r < R; r++) h->own_off[r + 1] = h->own_off[r] + h->own_count[r];
  h->my_n = h->own_count[G.rank];
  auto F = [](int32_t *&p) { if (p) devbuf_free(p); p = nullptr; };
  F(h->d_shuf_of_internal); F(h->d_dest_rank); F(h->d_dest_pos); F(h->d_my_internal);
  SQ_CHECK(devbuf_alloc((void **)&h->d_shuf_of_internal, n * sizeof(int32_t)));
  SQ_CHECK(devbuf_alloc((void **)&h->d_dest_rank, std::max<int64_t>(nloc, 1) * sizeof(int32_t)));
  SQ_CHECK(devbuf_alloc((void **)&h->d_dest_pos, std::max<int64_t>(nloc, 1) * sizeof(int32_t)));
  SQ_CHECK(devbuf_alloc((void **)&h->d_my_internal, std::max<int64_t>(h->my_n, 1) * sizeof(int32_t)));
  if (!h->d_scat_counter) {
    SQ_CUDA(cudaMalloc(&h->d_scat_counter, sizeof(unsigned long long)));
    SQ_CUDA(cudaMemset(h->d_scat_counter, 0, sizeof(unsigned long long)));
  }
  DevBuf<int32_t> owner, owner_sorted, rows, sorted_row, shuf_pos;
  SQ_CHECK(owner.alloc(n));
  SQ_CHECK(owner_sorted.alloc(n));
  SQ_CHECK(rows.alloc(n));
  SQ_CHECK(sorted_row.alloc(n));
  SQ_CHECK(shuf_pos.alloc(n));
  SQ_CUDA(cudaMemcpyAsync(owner.p, owner_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  liota_kernel<<<lblocks(n), 256, 0, s>>>(rows.p, n);
  SQ_LAUNCH_CHECK();
  // stable sort by owner: slice-concatenated order = rank 0's determinants in ascending caller index, then rank 1's, ...
  int bits = 1;
  while ((1 << bits) < R) bits++;
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, owner.p, owner_sorted.p, rows.p, sorted_row.p, (int)n, 0, bits, s);
  DevBuf<char> tmp;
  SQ_CHECK(tmp.alloc((int64_t)tb + 16));
  SQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, owner.p, owner_sorted.p, rows.p, sorted_row.p, (int)n, 0, bits, s));
  g_launch_count += 3;
  invert_kernel<<<lblocks(n), 256, 0, s>>>(sorted_row.p, shuf_pos.p, n);
  SQ_LAUNCH_CHECK();
  shuf_of_internal_kernel<<<lblocks(n), 256, 0, s>>>(h->d_perm, shuf_pos.p, h->d_shuf_of_internal, n);
  SQ_LAUNCH_CHECK();
  OffTable T;
  for (int r = 0; r < kMaxRanks; r++) T.off[r] = r < R ? h->own_off[r] : 0;
  if (nloc > 0) {
    dest_kernel<<<lblocks(nloc), 256, 0, s>>>(h->d_perm, owner.p, shuf_pos.p, T, h->row0, nloc, h->d_dest_rank, h->d_dest_pos);
    SQ_LAUNCH_CHECK();
  }
  if (h->my_n > 0) {
    my_internal_kernel<<<lblocks(h->my_n), 256, 0, s>>>(sorted_row.p, h->d_iperm, h->own_off[G.rank], h->my_n, h->d_my_internal);
    SQ_LAUNCH_CHECK();
  }
  SQ_CUDA(cudaStreamSynchronize(s));
  h->own_set = true;
  if (n_owned_out) *n_owned_out = h->my_n;
  return 0;
}

// allgatherv of the owned slices (slice-concatenated order) in place; NCCL fallback of the peer-store gather
static int allgather_slices(sqmc_b200_handle *h, double *buf, cudaStream_t s) {
  if (G.nranks == 1) return 0;
  ncclGroupStart();
  for (int r = 0; r < G.nranks; r++) {
    if (h->own_count[r] == 0) continue;
    ncclBroadcast(buf + h->own_off[r], buf + h->own_off[r], h->own_count[r], ncclDouble, r, G.comm, s);
  }
  ncclResult_t rc = ncclGroupEnd();
  if (rc != ncclSuccess) { set_error("allgather_slices: NCCL error %s", ncclGetErrorString(rc)); return 3; }
  return 0;
}

// host slice (my_n entries) -> the WHOLE vector in internal order in h->d_x, on every rank
int load_local_vector(sqmc_b200_handle *h, const double *host_slice, cudaStream_t s) {
  if (!h->own_set) { set_error("distributed-slice call before sqmc_b200_set_ownership"); return 2; }
  double *stage = h->d_tmp + h->own_off[G.rank];
  if (h->my_n > 0) SQ_CUDA(cudaMemcpyAsync(stage, host_slice, h->my_n * sizeof(double), cudaMemcpyHostToDevice, s));
  const double *xs = h->d_tmp;
  if (G.nranks > 1) {
    if (h->p2p.on) {
      double *f = nullptr;
      SQ_CHECK(p2p_gather(h, stage, h->my_n, h->own_off[G.rank], 1, s, &f));
      xs = f;
    } else {
      SQ_CHECK(allgather_slices(h, h->d_tmp, s));
    }
  }
  return permute_gather(xs, h->d_shuf_of_internal, h->d_x, h->n, s);  // x_int[p] = xs[shuf_of_internal[p]]
}

__global__ void gather_rows_kernel(const double *src, const int32_t *idx, double *dst, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// this rank's result block (internal order, nloc entries) -> the owners; host_slice receives the my_n owned entries.
// Synchronises the stream.
int store_local_vector(sqmc_b200_handle *h, const double *block, double *host_slice, cudaStream_t s) {
  if (!h->own_set) { set_error("distributed-slice call before sqmc_b200_set_ownership"); return 2; }
  const int64_t nloc = h->row1 - h->row0;
  const double *src = nullptr;
  if (G.nranks == 1) {
    SQ_CHECK(permute_scatter(block, h->d_dest_pos, h->d_tmp, nloc, s));  // d_tmp[caller row] = block[internal row]
    src = h->d_tmp;
  } else if (h->p2p.on) {
    double *y = nullptr;
    SQ_CHECK(p2p_to_owners(h, block, h->d_dest_rank, h->d_dest_pos, nloc, s, &y));
    src = y;
  } else {
    if (nloc > 0 && block != h->d_x + h->row0) SQ_CUDA(cudaMemcpyAsync(h->d_x + h->row0, block, nloc * sizeof(double), cudaMemcpyDeviceToDevice, s));
    SQ_CHECK(allgather_rows(h, h->d_x, s));
    if (h->my_n > 0) {
      gather_rows_kernel<<<lblocks(h->my_n), 256, 0, s>>>(h->d_x, h->d_my_internal, h->d_tmp, h->my_n);
      SQ_LAUNCH_CHECK();
    }
    src = h->d_tmp;
  }
  if (h->my_n > 0) SQ_CUDA(cudaMemcpyAsync(host_slice, src, h->my_n * sizeof(double), cudaMemcpyDeviceToHost, s));
  SQ_CUDA(cudaStreamSynchronize(s));
  return p2p_check(h, s);
}

// H.v (+ c*w) on the gathered vector in h->d_x, results delivered to their owners; host_slice receives this rank's slice.
// Row-bundle layout: ONE kernel multiplies and sends every row sum to its owner's buffer (bundle_hv_kernel<SCAT>);
// plain rows (SQMC_BUNDLE=0) or the NCCL fallback: multiply, then exchange.
static int multiply_to_owners(sqmc_b200_handle *h, double c, bool add_w, double *host_slice, cudaStream_t s) {
  const int64_t nloc = h->row1 - h->row0;
  const bool peers = G.nranks > 1 && h->p2p.on;
  if (h->bundle_R && (G.nranks == 1 || peers)) {
    OwnerScatter O = {};
    double *y = nullptr;
    if (peers) {
      SQ_CHECK(p2p_owner_begin(h, O, s));
    } else {  // one rank: the "owner buffer" is the caller-order staging vector
      O.nranks = 1;
      O.dst[0] = h->d_tmp;
      O.counter = &h->d_scat_counter[0];
      O.epoch = 1;
    }
    O.owner = h->d_dest_rank;
    O.pos = h->d_dest_pos;
    O.w = add_w ? h->d_x + h->row0 : nullptr;
    O.c = c;
    SQ_CHECK(bundle_spmv_scatter(h, h->d_x, O, s));
    if (peers) SQ_CHECK(p2p_owner_wait(h, O, s, &y));
    else y = h->d_tmp;
    if (h->my_n > 0) SQ_CUDA(cudaMemcpyAsync(host_slice, y, h->my_n * sizeof(double), cudaMemcpyDeviceToHost, s));
    SQ_CUDA(cudaStreamSynchronize(s));
    return p2p_check(h, s);
  }
  SQ_CHECK(spmv_launch(h, h->d_x, h->d_y, s));
  if (add_w) SQ_CHECK(projector_epilogue(h->d_y, h->d_x + h->row0, c, nloc, s));
  return store_local_vector(h, h->d_y, host_slice, s);
}

int matvec_local(sqmc_b200_handle *h, const double *x_local, double *y_local) {
  cudaStream_t s = G.stream;
  SQ_CHECK(load_local_vector(h, x_local, s));
  return multiply_to_owners(h, 0.0, false, y_local, s);
}

// deltaw = Hstored . w (do_walk.f90:2259) + e_trial*tau*w (:2290), reduce-scattered to the owners (:2260)
int projector_local(sqmc_b200_handle *h, double tau, double e_trial, const double *w_local, double *deltaw_local) {
  cudaStream_t s = G.stream;
  SQ_CHECK(load_local_vector(h, w_local, s));
  return multiply_to_owners(h, e_trial * tau, true, deltaw_local, s);
}

}  // namespace sqmc
