#!/usr/bin/env python
"""In-process A/B of build variants on the same determinant space: build_ab.py [n_dets] [space] ENV=val,ENV=val ..."""
import os, sys, time, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sqmc_b200 as sq
from sqmc_b200 import _lib, spaces
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
space = sys.argv[2] if len(sys.argv) > 2 else "hci"
variants = sys.argv[3:] or ["SQMC_CONNECT_STAGE=0", "SQMC_CONNECT_STAGE=1", "SQMC_CONNECT_STAGE=0", "SQMC_CONNECT_STAGE=1"]
_lib.init(device=0)
chem = sq.ChemSystem("data/C2_v2z_curve/r1.24253/FCIDUMP")
H = sq.SparseHamiltonian(chem)
if space == "hci":
    up, dn, _, _ = spaces.hci_space(H, chem, n)
else:
    up, dn, _ = spaces.c2_lowest_energy_space(chem, n)
x = spaces.splitmix_vector(len(up))
ref = None
for v in variants:
    for kv in v.split(","):
        k, val = kv.split("=")
        os.environ[k] = val
    t0 = time.perf_counter()
    nnz = H.generate_sparse_ham_upper_triangular(up, dn)
    wall = time.perf_counter() - t0
    y = H.fast_sparse_matrix_multiply_upper_triangular(x)
    sig = hashlib.sha256(y.tobytes()).hexdigest()[:16]
    if ref is None: ref = sig
    print(json.dumps({"variant": v, "n": len(up), "nnz_upper": int(nnz), "wall_s": wall, "phases": H.build_times(), "y_sha": sig, "same_as_first": sig == ref}), flush=True)
