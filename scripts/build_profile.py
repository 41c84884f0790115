#!/usr/bin/env python
"""Host wall-clock profile of the H build (SQMC_BUILD_PROFILE=1): builds the same 10^7-determinant matrix three times."""
import os, sys, time
os.environ["SQMC_BUILD_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqmc_b200 as sq
from sqmc_b200 import _lib, spaces
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
_lib.init(device=0)
chem = sq.ChemSystem("data/C2_v2z_curve/r1.24253/FCIDUMP")
up, dn, _ = spaces.c2_lowest_energy_space(chem, n)
H = sq.SparseHamiltonian(chem)
for k in range(3):
    t0 = time.perf_counter()
    nnz = H.generate_sparse_ham_upper_triangular(up, dn)
    print("build %d: %.3f s wall, nnz_upper %d, phases %s" % (k, time.perf_counter() - t0, nnz, H.build_times()), file=sys.stderr, flush=True)
