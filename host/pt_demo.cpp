// pt_demo.cpp -- a compiled caller of the second-order PT entry points, doing what the semistochastic branch of do_pt
// (hci.f90:4245-4300) does: the deterministic correction with eps_pt_big (second_order_pt, hci.f90:1100), then the stochastic
// difference between eps_pt and eps_pt_big (second_order_pt_alias, hci.f90:1314) with the caller's rannyu state, through the
// C++ host mirror (host/sqmc_b200_host.hpp).  Input / output are flat binary files written / read by tests/test_gpu_host_cpp.py.
//   usage: pt_demo <input.bin> <output.bin>
#include <cstdio>
#include <cstdlib>

#include "sqmc_b200_host.hpp"

using namespace sqmc_b200_host;

template <typename T>
static void rd(FILE *f, std::vector<T> &v, size_t n) {
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]); return 2; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror("input"); return 2; }
  std::vector<int64_t> hdr;
  rd(f, hdr, 12);  // norb n_dim nup ndn n n_mc max_samples seed[4] unused
  const int norb = (int)hdr[0], n_dim = (int)hdr[1], nup = (int)hdr[2], ndn = (int)hdr[3], n_mc = (int)hdr[5], max_samples = (int)hdr[6];
  const int64_t n = hdr[4];
  int32_t irand_state[4] = {(int32_t)hdr[7], (int32_t)hdr[8], (int32_t)hdr[9], (int32_t)hdr[10]};
  std::vector<rk> par, kvec, wts;
  rd(f, par, 5);  // length_cell var_energy eps_pt eps_pt_big target_error
  rd(f, kvec, (size_t)norb * n_dim);
  std::vector<ik> up, dn;
  rd(f, up, n);
  rd(f, dn, n);
  rd(f, wts, n);
  fclose(f);
  try {
    model_system S = model_system::heg(norb, n_dim, kvec, par[0], nup, ndn);
    rk pt_big = 0, pt_diff = 0, pt_diff_std_dev = 0;
    int64_t nconn_big = 0, nconn = 0;
    int n_samples = 0;
    second_order_pt(S, n, up.data(), dn.data(), wts, par[1], par[3], pt_big, nconn_big);
    second_order_pt_alias(S, n, up.data(), dn.data(), wts, par[1], par[2], n_mc, par[4], par[3], irand_state, pt_diff, pt_diff_std_dev, nconn, n_samples,
                          max_samples);
    FILE *o = fopen(argv[2], "wb");
    const double out[7] = {pt_big, (double)nconn_big, pt_diff, pt_diff_std_dev, (double)n_samples, (double)nconn, (double)irand_state[3]};
    fwrite(out, sizeof(double), 7, o);
    fclose(o);
    printf("pt_demo: Second-order PT energy lowering= %15.9f +- %11.9f (%13.9f %12.9f), %d samples\n", pt_big + pt_diff, pt_diff_std_dev, pt_big, pt_diff,
           n_samples);
  } catch (const std::exception &e) {
    fprintf(stderr, "pt_demo failed: %s\n", e.what());
    return 1;
  }
  sqmc_b200_finalize();
  return 0;
}
