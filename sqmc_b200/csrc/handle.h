// handle.h -- opaque handle behind the C ABI (include/sqmc_b200.h) and the
// process-wide device / communicator state.
#pragma once
#include <nccl.h>

#include <vector>

#include "common.cuh"
#include "elements.cuh"

namespace sqmc {

struct Global {
  bool inited = false;
  int device = 0, rank = 0, nranks = 1;
  int sm_count = 148;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;  // all library work is issued on this stream unless a stream is passed in
};
extern Global G;

// number of degree bins for the SpMV (sub-warp vector sizes 2,4,8,16,32 + CTA-per-row)
static const int kNumBins = 6;

}  // namespace sqmc

struct sqmc_b200_handle {
  sqmc::ModelTables T;  // device pointers inside
  int NW = 1;
  // owned device tables
  double *d_integrals = nullptr;
  int32_t *d_combine_2 = nullptr;
  double *d_kvec = nullptr;
  int32_t *d_hkvec = nullptr;
  double *d_kenergies = nullptr;

  // ---- determinant list (internal = alpha-major sorted order) ----
  int64_t n = 0;
  uint64_t *d_up = nullptr, *d_dn = nullptr;  // n*NW words each, internal order
  int32_t *d_perm = nullptr;                  // internal row -> caller row (0-based)
  int32_t *d_iperm = nullptr;                 // caller row -> internal row

  // ---- row sharding (contiguous internal row blocks, balanced by nnz) ----
  std::vector<int64_t> row_starts;  // nranks+1
  int64_t row0 = 0, row1 = 0;       // this rank's rows [row0,row1)

  // ---- full symmetric CSR of the local rows, columns in internal numbering ----
  int64_t *d_rowptr = nullptr;  // (row1-row0+1) offsets into cols/vals
  int32_t *d_cols = nullptr;
  double *d_vals = nullptr;
  int64_t nnz_local = 0, capacity = 0;
  int64_t nnz_full = 0, nnz_upper = 0;  // global
  double scale = 1.0;                   // product of scale_values() ratios applied to d_vals

  // ---- SpMV degree bins: row lists (local row ids) ----
  int32_t *d_bin_rows = nullptr;  // concatenated lists
  int64_t bin_off[sqmc::kNumBins + 1] = {0};

  // ---- work buffers ----
  double *d_x = nullptr;   // n (global length, internal order)
  double *d_y = nullptr;   // local rows
  double *d_tmp = nullptr; // n (caller order staging)
  double build_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // [5] candidates (local) [6] alpha groups [7] beta groups
};

namespace sqmc {
// build.cu
int build_h(sqmc_b200_handle *h, int64_t n, const void *dets_up, const void *dets_dn, int64_t ndet_old);
int diagonal(sqmc_b200_handle *h, int64_t n, const void *dets_up, const void *dets_dn, double *diag);
int export_upper(sqmc_b200_handle *h, int64_t *counts, int64_t *indices, double *values);
int import_upper(sqmc_b200_handle *h, int64_t n, const int64_t *counts, const int64_t *indices, const double *values);
void free_matrix(sqmc_b200_handle *h);
void partition_rows(const int64_t *prefix, int64_t n, int nranks, int64_t *starts);
// spmv.cu
int spmv_setup_bins(sqmc_b200_handle *h);
int spmv_launch(sqmc_b200_handle *h, const double *x_dev, double *y_dev, cudaStream_t s);
int permute_gather(const double *src, const int32_t *idx, double *dst, int64_t n, cudaStream_t s);   // dst[i] = src[idx[i]]
int permute_scatter(const double *src, const int32_t *idx, double *dst, int64_t n, cudaStream_t s);  // dst[idx[i]] = src[i]
int scale_array(double *a, int64_t n, double r, cudaStream_t s);
int projector_epilogue(double *deltaw, const double *w, double c, int64_t n, cudaStream_t s);  // deltaw += c*w
int allgather_rows(sqmc_b200_handle *h, double *x_full, cudaStream_t s);  // in-place allgather of row blocks
// davidson.cu
int davidson(sqmc_b200_handle *h, int n_states, const double *v0, double *evecs, double *evals, double tol, int max_vec,
             int *n_matvec_out, double *ritz_log, int ritz_log_cap, int *n_ritz_logged);
}  // namespace sqmc
