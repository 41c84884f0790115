// sqmc_b200_host.hpp -- C++ host-side mirror of the reference's Fortran interface for the hot path, written above the
// C ABI of libsqmc_b200.so (include/sqmc_b200.h).  The reference is compiled Fortran and this image has no Fortran
// compiler, so this header plays the role of fortran/sqmc_b200_iface.f90 for compiled callers: the routines keep the
// reference's names, argument order and meaning (allocatable intent(out) arrays become std::vector&), and errors throw
// (the reference `stop`s).  One `model_system` object stands for the module globals the Fortran routines read.
//
//   generate_sparse_ham_chem_upper_triangular      chemistry.f90:7639
//   generate_sparse_ham_heg_upper_triangular       heg.f90:3553
//   generate_sparse_ham_hubbardk_upper_triangular  hubbard.f90:9435
//   fast_sparse_matrix_multiply_upper_triangular   more_tools.f90:3622
//   davidson_sparse                                more_tools.f90:2018
//   deterministic projector step                   do_walk.f90:2255-2325
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/sqmc_b200.h"

namespace sqmc_b200_host {

typedef unsigned __int128 ik;  // integer(ik), types.f90:26
typedef int64_t i8b;           // types.f90:18
typedef double rk;             // types.f90:27

inline void check(int rc) {
  if (rc != 0) throw std::runtime_error(std::string("sqmc_b200: ") + sqmc_b200_last_error());
}

// sparse_mat of commons/common_selected_ci.f90:26-31: only ndet is consulted (rows already built)
struct sparse_mat {
  i8b ndet = 0;
};

// stands for the module globals of chemistry.f90 / heg.f90 / hubbard.f90 that the builders read
class model_system {
 public:
  sqmc_b200_handle *h = nullptr;
  i8b n = 0;
  static void init(int device = 0) {
    static bool done = false;
    if (!done) check(sqmc_b200_init(device, 0, 1, nullptr));
    done = true;
  }
  static model_system chem(int norb, int nup, int ndn, const std::vector<rk> &integrals, const std::vector<int32_t> &combine_2, bool time_sym, int z) {
    model_system s;
    init();
    check(sqmc_b200_system_chem(&s.h, norb, nup, ndn, integrals.data(), (i8b)integrals.size(), combine_2.data(), time_sym ? 1 : 0, z));
    return s;
  }
  static model_system heg(int norb, int n_dim, const std::vector<rk> &k_vectors, rk length_cell, int nup, int ndn) {
    model_system s;
    init();
    check(sqmc_b200_system_heg(&s.h, norb, n_dim, k_vectors.data(), length_cell, nup, ndn));
    return s;
  }
  static model_system hubbardk(int l_x, int l_y, const std::vector<int32_t> &k_vectors, const std::vector<rk> &k_energies, rk ubyn, int nup, int ndn) {
    model_system s;
    init();
    check(sqmc_b200_system_hubbardk(&s.h, l_x, l_y, k_vectors.data(), k_energies.data(), ubyn, nup, ndn));
    return s;
  }
  model_system() = default;
  model_system(model_system &&o) noexcept : h(o.h), n(o.n) { o.h = nullptr; }
  model_system &operator=(model_system &&o) noexcept {
    if (this != &o) { release(); h = o.h; n = o.n; o.h = nullptr; }
    return *this;
  }
  model_system(const model_system &) = delete;
  model_system &operator=(const model_system &) = delete;
  ~model_system() { release(); }

 private:
  void release() {
    if (h) sqmc_b200_free(h);
    h = nullptr;
  }
};

namespace detail {
inline void build(model_system &S, const std::vector<ik> &dets_up, const std::vector<ik> &dets_dn, std::vector<i8b> &H_indices,
                  std::vector<i8b> &H_nonzero_elements, std::vector<rk> &H_values, bool hf_to_psit, sparse_mat *sparse_ham, rk *average_connections) {
  if (dets_up.size() != dets_dn.size() || dets_up.empty()) throw std::runtime_error("sqmc_b200: bad determinant list");
  const i8b n = (i8b)dets_up.size();
  i8b nnz = 0;
  check(sqmc_b200_set_hf_to_psit(S.h, hf_to_psit ? 1 : 0));
  check(sqmc_b200_build_h(S.h, n, dets_up.data(), dets_dn.data(), sparse_ham ? sparse_ham->ndet : 0, &nnz));
  S.n = n;
  H_nonzero_elements.assign(n, 0);
  H_indices.assign(nnz, 0);
  H_values.assign(nnz, 0.0);
  check(sqmc_b200_export_upper(S.h, H_nonzero_elements.data(), H_indices.data(), H_values.data()));
  if (sparse_ham) sparse_ham->ndet = n;                       // chemistry.f90:7989
  if (average_connections) *average_connections = (rk)nnz / n;  // chemistry.f90:8006
}
}  // namespace detail

// chemistry.f90:7639.  time_sym_on is carried by the model_system (module variable time_sym); hf_to_psit is ignored by the
// reference's partial-connection branch (:7721) and therefore here.
inline void generate_sparse_ham_chem_upper_triangular(model_system &S, const std::vector<ik> &dets_up, const std::vector<ik> &dets_dn,
                                                      std::vector<i8b> &H_indices, std::vector<i8b> &H_nonzero_elements, std::vector<rk> &H_values,
                                                      bool /*time_sym_on*/, bool /*hf_to_psit*/, sparse_mat *sparse_ham = nullptr,
                                                      rk *average_connections = nullptr) {
  detail::build(S, dets_up, dets_dn, H_indices, H_nonzero_elements, H_values, false, sparse_ham, average_connections);
}
// heg.f90:3553
inline void generate_sparse_ham_heg_upper_triangular(model_system &S, const std::vector<ik> &dets_up, const std::vector<ik> &dets_dn,
                                                     std::vector<i8b> &H_indices, std::vector<i8b> &H_nonzero_elements, std::vector<rk> &H_values,
                                                     bool /*time_sym_on*/, bool /*hf_to_psit*/, sparse_mat *sparse_ham = nullptr) {
  detail::build(S, dets_up, dets_dn, H_indices, H_nonzero_elements, H_values, false, sparse_ham, nullptr);
}
// hubbard.f90:9435 (hf_to_psit honoured: row 1 = single zero diagonal entry, :9636-9643)
inline void generate_sparse_ham_hubbardk_upper_triangular(model_system &S, const std::vector<ik> &dets_up, const std::vector<ik> &dets_dn,
                                                          std::vector<i8b> &H_indices, std::vector<i8b> &H_nonzero_elements,
                                                          std::vector<rk> &H_values, bool hf_to_psit) {
  detail::build(S, dets_up, dets_dn, H_indices, H_nonzero_elements, H_values, hf_to_psit, nullptr, nullptr);
}

// more_tools.f90:3622: answer = H . vector with the matrix the last generate_sparse_ham_* left resident on the device
// (the reference passes matrix_indices / nelem_nonzero / matrix_values explicitly; they are the arrays returned above)
inline void fast_sparse_matrix_multiply_upper_triangular(model_system &S, int n, const std::vector<rk> &vector, std::vector<rk> &answer) {
  if ((i8b)n != S.n || (i8b)vector.size() < S.n) throw std::runtime_error("sqmc_b200: matvec size mismatch");
  answer.assign(n, 0.0);
  check(sqmc_b200_matvec(S.h, vector.data(), answer.data(), 1, n));
}

// more_tools.f90:2018: final_vector(n, n_states) column-major, lowest_eigenvalues(n_states); initial_vector optional
inline void davidson_sparse(model_system &S, int n, int n_states, std::vector<rk> &final_vector, std::vector<rk> &lowest_eigenvalues,
                            const std::vector<rk> *initial_vector = nullptr, std::vector<rk> *iteration_eigenvalues = nullptr) {
  if ((i8b)n != S.n) throw std::runtime_error("sqmc_b200: davidson size mismatch");
  final_vector.assign((size_t)n * n_states, 0.0);
  lowest_eigenvalues.assign(n_states, 0.0);
  std::vector<rk> log(1024 * (size_t)n_states);
  int nmv = 0, nlog = 0;
  check(sqmc_b200_davidson(S.h, n_states, initial_vector ? initial_vector->data() : nullptr, final_vector.data(), lowest_eigenvalues.data(), 1.e-10,
                           50, &nmv, log.data(), 1024, &nlog));
  if (iteration_eigenvalues) iteration_eigenvalues->assign(log.begin(), log.begin() + (size_t)std::min(nlog, 1024) * n_states);
}

// do_walk.f90:2259-2290 with the stored matrix already scaled by -tau (semistoch.f90:657,880)
inline void scale_values(model_system &S, rk ratio) { check(sqmc_b200_scale_values(S.h, ratio)); }
// matrix_lanczos_sparse(n, lowest_eigenvector, lowest_eigenvalue, matrix_indices, nelem_nonzero, matrix_values,
//                       highest_eigenvalue, second_lowest_eigenvalue, initial_vector)   (more_tools.f90:1742)
// the matrix arguments are the handle's resident matrix; optional outputs / input as pointers (null = absent)
inline void matrix_lanczos_sparse(model_system &S, std::vector<rk> &lowest_eigenvector, rk &lowest_eigenvalue, rk *highest_eigenvalue = nullptr,
                                  rk *second_lowest_eigenvalue = nullptr, const std::vector<rk> *initial_vector = nullptr) {
  int64_t n = 0, nu = 0, nf = 0;
  check(sqmc_b200_nnz(S.h, &n, &nu, &nf));
  lowest_eigenvector.assign((size_t)n, 0.0);
  rk eig3[3] = {0, 0, 0};
  int nit = 0, nlog = 0;
  check(sqmc_b200_lanczos(S.h, initial_vector ? initial_vector->data() : nullptr, lowest_eigenvector.data(), eig3, 1.0e-10, 50, &nit, nullptr, 0, &nlog));
  lowest_eigenvalue = eig3[0];
  if (highest_eigenvalue) *highest_eigenvalue = eig3[1];
  if (second_lowest_eigenvalue) *second_lowest_eigenvalue = eig3[2];
}
// davidson_sparse_single(n, lowest_eigenvector, lowest_eigenvalue, matrix_indices, nelem_nonzero, matrix_values,
//                        highest_eigenvalue, initial_vector)   (more_tools.f90:3055)
inline void davidson_sparse_single(model_system &S, std::vector<rk> &lowest_eigenvector, rk &lowest_eigenvalue, rk *highest_eigenvalue = nullptr,
                                   const std::vector<rk> *initial_vector = nullptr) {
  int64_t n = 0, nu = 0, nf = 0;
  check(sqmc_b200_nnz(S.h, &n, &nu, &nf));
  lowest_eigenvector.assign((size_t)n, 0.0);
  rk eig2[2] = {0, 0};
  int nit = 0, nlog = 0;
  check(sqmc_b200_davidson_single(S.h, initial_vector ? initial_vector->data() : nullptr, lowest_eigenvector.data(), eig2, 1.0e-10, 50, &nit, nullptr, 0, &nlog));
  lowest_eigenvalue = eig2[0];
  if (highest_eigenvalue) *highest_eigenvalue = eig2[1];
}
// second_order_pt(ndets, dets_up, dets_dn, wts, diag_elems, var_energy, eps_pt, delta_e_2pt, ndets_connected)   (hci.f90:1100)
// diag_elems is not needed (H_aa is evaluated on the device); determinants as 16-byte integers like the reference's integer(ik)
inline void second_order_pt(model_system &S, int64_t ndets, const void *dets_up, const void *dets_dn, const std::vector<rk> &wts, rk var_energy,
                            rk eps_pt, rk &delta_e_2pt, int64_t &ndets_connected) {
  check(sqmc_b200_pt2(S.h, ndets, dets_up, dets_dn, wts.data(), var_energy, eps_pt, &delta_e_2pt, &ndets_connected));
}
// second_order_pt_alias(ndets, dets_up, dets_dn, wts, diag_elems, var_energy, eps_pt, n_mc, target_error, eps_pt_big, pt_energy,
//                       pt_energy_std_dev, ndets_connected, pt_big)   (hci.f90:1314), n_mc > 0.  irand_state = savern(): the advanced
// state comes back for setrn, so the caller's rannyu stream continues as after the reference routine.
inline void second_order_pt_alias(model_system &S, int64_t ndets, const void *dets_up, const void *dets_dn, const std::vector<rk> &wts, rk var_energy,
                                  rk eps_pt, int n_mc, rk target_error, rk eps_pt_big, int32_t irand_state[4], rk &pt_energy, rk &pt_energy_std_dev,
                                  int64_t &ndets_connected, int &n_samples, int max_samples = 1000000) {
  check(sqmc_b200_pt2_alias(S.h, ndets, dets_up, dets_dn, wts.data(), var_energy, eps_pt, eps_pt_big, n_mc, target_error, irand_state, max_samples,
                            &pt_energy, &pt_energy_std_dev, &n_samples, nullptr, &ndets_connected));
}
// ---- the reference's MPI data distribution: every rank passes / receives the slice of the determinants it owns ----
// set once per determinant list: owner_of_row(i) = get_det_owner(dets_up(i), dets_dn(i)) (mpi_routines.f90:419); returns my_nimp
inline int64_t set_ownership(model_system &S, const std::vector<int32_t> &owner_of_row) {
  if ((i8b)owner_of_row.size() != S.n) throw std::runtime_error("sqmc_b200: owner_of_row must have n entries");
  int64_t mine = 0;
  check(sqmc_b200_set_ownership(S.h, owner_of_row.data(), &mine));
  return mine;
}
// fast_sparse_matrix_multiply_local_band(n_imp, ..., vector = walk_wt(my_locations_of_imp_dets(1:my_nimp)), answer = deltaw)
// followed by mpi_redscatt_real_dparray(deltaw(1:n_imp), imp_core_mask)   (do_walk.f90:2259-2260, more_tools.f90:3562, mpi_routines.f90:1592):
// on return answer(1:my_nimp) holds this rank's rows of H . vector
inline void fast_sparse_matrix_multiply_local_band(model_system &S, const std::vector<rk> &vector_local, std::vector<rk> &answer_local) {
  answer_local.assign(vector_local.size(), 0.0);
  check(sqmc_b200_matvec_local(S.h, vector_local.data(), answer_local.data(), 1, (i8b)std::max<size_t>(vector_local.size(), 1)));
}
// the same + e_trial*tau*my_imp_wt (do_walk.f90:2290): deltaw(1:my_nimp)
inline void deterministic_projector_step_local(model_system &S, rk tau, rk e_trial, const std::vector<rk> &my_imp_wt, std::vector<rk> &deltaw_local) {
  deltaw_local.assign(my_imp_wt.size(), 0.0);
  check(sqmc_b200_projector_local(S.h, tau, e_trial, my_imp_wt.data(), deltaw_local.data()));
}
// davidson_sparse_mpi2(remote_det_map, local_det_map, n_states, final_vector, lowest_eigenvalues, ..., initial_vector)
// (more_tools.f90:2525): vectors are local_det_map%ndets x n_states
inline void davidson_sparse_mpi2(model_system &S, int64_t ndets_local, int n_states, std::vector<rk> &final_vector, std::vector<rk> &lowest_eigenvalues,
                                 const std::vector<rk> *initial_vector = nullptr) {
  final_vector.assign((size_t)std::max<int64_t>(ndets_local, 1) * n_states, 0.0);
  lowest_eigenvalues.assign(n_states, 0.0);
  int nmv = 0, nlog = 0;
  check(sqmc_b200_davidson_local(S.h, n_states, initial_vector ? initial_vector->data() : nullptr, final_vector.data(), lowest_eigenvalues.data(), 1.e-10, 50,
                                 &nmv, nullptr, 0, &nlog));
}
// storage-order hint for H.v (0 = plain rows, 2/4 = column-merged bundles); no reference counterpart
inline void set_row_bundle(model_system &S, int rows_per_bundle) { check(sqmc_b200_set_row_bundle(S.h, rows_per_bundle)); }
inline void deterministic_projector_step(model_system &S, rk tau, rk e_trial, const std::vector<rk> &imp_wt, std::vector<rk> &deltaw) {
  deltaw.assign(S.n, 0.0);
  check(sqmc_b200_projector(S.h, tau, e_trial, imp_wt.data(), deltaw.data()));
}

}  // namespace sqmc_b200_host
