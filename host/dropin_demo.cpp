// dropin_demo.cpp -- a compiled caller of the hot path, doing what diagonalize_sparse_hamiltonian_chem_excited
// (chemistry.f90:6057-6182) and the deterministic-projector part of walk (do_walk.f90:2255-2325) do, through the C++
// host mirror (host/sqmc_b200_host.hpp) of the reference's Fortran interface.  Input / output are flat binary files
// written / read by tests/test_gpu_host_cpp.py.
//   usage: dropin_demo <input.bin> <output.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sqmc_b200_host.hpp"

using namespace sqmc_b200_host;

template <typename T>
static void rd(FILE *f, std::vector<T> &v, size_t n) {
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}
template <typename T>
static void wr(FILE *f, const std::vector<T> &v) {
  int64_t n = (int64_t)v.size();
  fwrite(&n, sizeof n, 1, f);
  if (n) fwrite(v.data(), sizeof(T), n, f);
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]); return 2; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror("input"); return 2; }
  std::vector<int64_t> hdr;
  rd(f, hdr, 12);  // model norb nup ndn time_sym z a b nint n n_states hf_to_psit
  const int model = (int)hdr[0], norb = (int)hdr[1], nup = (int)hdr[2], ndn = (int)hdr[3], time_sym = (int)hdr[4], z = (int)hdr[5];
  const int64_t nint = hdr[8], n = hdr[9];
  const int n_states = (int)hdr[10], hf_to_psit = (int)hdr[11];
  try {
    model_system S;
    if (model == 0) {
      std::vector<rk> integrals;
      std::vector<int32_t> c2;
      rd(f, integrals, nint);
      rd(f, c2, (size_t)(norb + 1) * (norb + 1));
      S = model_system::chem(norb, nup, ndn, integrals, c2, time_sym != 0, z);
    } else {
      std::vector<int32_t> kv;
      std::vector<rk> ke, ubyn;
      rd(f, kv, (size_t)2 * norb);
      rd(f, ke, norb);
      rd(f, ubyn, 1);
      S = model_system::hubbardk((int)hdr[6], (int)hdr[7], kv, ke, ubyn[0], nup, ndn);
    }
    std::vector<ik> up, dn;
    rd(f, up, n);
    rd(f, dn, n);
    std::vector<rk> x;
    rd(f, x, n);
    fclose(f);

    std::vector<i8b> H_indices, H_nonzero_elements;
    std::vector<rk> H_values, answer, final_vector, lowest_eigenvalues, ritz, deltaw;
    sparse_mat sparse_ham;
    rk average_connections = 0;
    if (model == 0) generate_sparse_ham_chem_upper_triangular(S, up, dn, H_indices, H_nonzero_elements, H_values, time_sym != 0, false, &sparse_ham, &average_connections);
    else generate_sparse_ham_hubbardk_upper_triangular(S, up, dn, H_indices, H_nonzero_elements, H_values, hf_to_psit != 0);
    fast_sparse_matrix_multiply_upper_triangular(S, (int)n, x, answer);
    davidson_sparse(S, (int)n, n_states, final_vector, lowest_eigenvalues, nullptr, &ritz);
    const rk tau = 0.01, e_trial = lowest_eigenvalues[0];
    scale_values(S, -tau);
    deterministic_projector_step(S, tau, e_trial, x, deltaw);

    FILE *o = fopen(argv[2], "wb");
    wr(o, H_nonzero_elements);
    wr(o, H_indices);
    wr(o, H_values);
    wr(o, answer);
    wr(o, lowest_eigenvalues);
    wr(o, final_vector);
    wr(o, ritz);
    wr(o, deltaw);
    fclose(o);
    printf("dropin_demo: n=%lld nnz=%lld average_connections=%.2f E0=%.10f\n", (long long)n, (long long)H_indices.size(), average_connections,
           lowest_eigenvalues[0]);
  } catch (const std::exception &e) {
    fprintf(stderr, "dropin_demo failed: %s\n", e.what());
    return 1;
  }
  sqmc_b200_finalize();
  return 0;
}
