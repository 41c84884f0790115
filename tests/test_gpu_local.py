"""Distributed-slice entry points (sqmc_b200_set_ownership / matvec_local / projector_local / davidson_local) on one GPU:
with a single rank the owned slice is the whole vector, so the calls must agree with the full-vector entry points and
with the oracle.  The N > 1 data movement is covered by tests/run_multi_gpu_parity.py (tests/test_gpu_multi.py)."""
import numpy as np
import pytest

import sqmc_b200 as sq

pytestmark = pytest.mark.gpu


def test_local_entry_points_single_rank(oracle, heg_space):
    S, r = heg_space
    cnt, idx, val = S.build_upper(r["up"], r["dn"])
    n = len(cnt)
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49))
    H.generate_sparse_ham_upper_triangular(r["up"], r["dn"])
    with pytest.raises(sq.SqmcError):
        H.matvec_local(np.zeros(n))                      # ownership not set yet
    with pytest.raises(sq.SqmcError):
        H.set_ownership(np.ones(n, dtype=np.int32))      # rank 1 does not exist in a one-rank job
    assert H.set_ownership(oracle.det_owner(r["up"], r["dn"], 1)) == n
    assert H.exchange_mode() == "single"
    x = np.random.default_rng(5).uniform(-1, 1, n)
    yref = oracle.matvec_upper(cnt, idx, val, x)
    y = H.matvec_local(x)
    assert np.max(np.abs(y - yref)) <= 1e-12 * np.max(np.abs(yref))
    assert np.array_equal(y, H.matvec(x))                # same kernels, same order of operations
    ref = oracle.davidson(cnt, idx, val, n_states=1)
    got = H.davidson_sparse_local(n_states=1)
    assert got["ritz"].shape == ref["ritz"].shape and np.max(np.abs(got["ritz"] - ref["ritz"])) < 1e-8
    full = H.davidson_sparse(n_states=1)
    assert np.array_equal(got["evecs"], full["evecs"])
    tau, e_trial = 0.01, float(val[0])
    H.scale_values(-tau)
    w = x / np.linalg.norm(x)
    H.register_host(w)
    dw = H.projector_step_local(tau, e_trial, w)
    H.unregister_host(w)
    _, dwr = oracle.projector_step(cnt, idx, -tau * val, tau, e_trial, w)
    assert np.max(np.abs(dw - dwr)) <= 1e-12 * np.max(np.abs(dwr)) + 1e-15
    # a rebuild drops the map: it belongs to the determinant list
    H.generate_sparse_ham_upper_triangular(r["up"][:277], r["dn"][:277])
    with pytest.raises(sq.SqmcError):
        H.matvec_local(np.zeros(277))
    H.close()
