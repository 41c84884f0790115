"""matrix_lanczos_sparse on the device (csrc/davidson.cu: lanczos) against the oracle's restatement of
more_tools.f90:1742-1883: per-step eigenvalues, the three returned eigenvalues and the eigenvector."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(got, ref):
    assert got["n_iter"] == ref["n_iter"]
    assert len(got["ritz"]) == len(ref["ritz"]) and np.max(np.abs(got["ritz"] - ref["ritz"])) < 1e-8
    assert abs(got["lowest_eigenvalue"] - ref["lowest"]) < 1e-8
    assert abs(got["highest_eigenvalue"] - ref["highest"]) < 1e-7
    assert abs(got["second_lowest_eigenvalue"] - ref["second_lowest"]) < 1e-7
    a, b = got["lowest_eigenvector"], ref["evec"]
    assert min(np.max(np.abs(a - b)), np.max(np.abs(a + b))) < 1e-6


def test_lanczos_hubbard_sector(oracle):
    """the reference's user of this solver: 4x4 Hubbard in k-space, U/t = 4, lowest 3000 dets of the k = 0 sector"""
    import sqmc_b200 as sq
    from sqmc_b200 import spaces
    hub = sq.HubbardKSystem(4, 4, 1.0, 4.0, 8, 8)
    up, dn, _ = spaces.hubbard_momentum_sector(hub, 3000)
    S = oracle.System.hubbardk(4, 4, 1.0, 4.0, 8, 8)
    cnt, idx, val = S.build_upper(up, dn)
    H = sq.SparseHamiltonian(hub)
    assert H.generate_sparse_ham_upper_triangular(up, dn) == len(idx)
    _check(H.matrix_lanczos_sparse(), oracle.lanczos(cnt, idx, val))
    v0 = np.cos(np.arange(len(up)) * 0.37) + 0.1
    _check(H.matrix_lanczos_sparse(initial_vector=v0), oracle.lanczos(cnt, idx, val, v0=v0))


def test_lanczos_heg_golden_energy(oracle, heg_space):
    """HEG 277-determinant space of the reference's e2e test: Lanczos converges to the golden Davidson energy"""
    import sqmc_b200 as sq
    s, r = heg_space
    up, dn = r["up"][:277], r["dn"][:277]
    cnt, idx, val = s.build_upper(up, dn)
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49))
    H.generate_sparse_ham_upper_triangular(up, dn)
    got = H.matrix_lanczos_sparse()
    _check(got, oracle.lanczos(cnt, idx, val))
    assert abs(got["lowest_eigenvalue"] - 58.2825967049) < 5e-9      # src/e2e_tests/heg/o_det_ref:270 (davidson_sparse, same matrix)


def test_davidson_sparse_single(oracle, heg_space, c2_space):
    """davidson_sparse_single (more_tools.f90:3055-3233): printed eigenvalues, lowest / highest eigenvalue, eigenvector"""
    import sqmc_b200 as sq
    from conftest import C2_FCIDUMP
    s, r = heg_space
    up, dn = r["up"][:277], r["dn"][:277]
    cnt, idx, val = s.build_upper(up, dn)
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49))
    H.generate_sparse_ham_upper_triangular(up, dn)
    # (a start vector far from the ground state makes this solver -- one guarded element only -- divide by nearly
    #  vanishing E - H_jj, and its late iterations then amplify rounding differences: a perturbed HF vector is used)
    for v0 in (None, np.eye(277)[0] + 0.01 * np.cos(np.arange(277) * 0.11)):
        got, ref = H.davidson_sparse_single(initial_vector=v0), oracle.davidson_single(cnt, idx, val, v0=v0)
        assert got["n_iter"] == ref["n_iter"] and len(got["ritz"]) == len(ref["ritz"])
        assert np.max(np.abs(got["ritz"] - ref["ritz"])) < 1e-8
        assert abs(got["lowest_eigenvalue"] - ref["lowest"]) < 1e-8 and abs(got["highest_eigenvalue"] - ref["highest"]) < 1e-8
        a, b = got["lowest_eigenvector"], ref["evec"]
        assert min(np.max(np.abs(a - b)), np.max(np.abs(a + b))) < 1e-6
    assert abs(H.davidson_sparse_single()["lowest_eigenvalue"] - 58.2825967049) < 5e-9     # o_det_ref:270
    s2, r2 = c2_space
    cnt, idx, val = s2.build_upper(r2["up"], r2["dn"])
    G = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP))
    G.generate_sparse_ham_upper_triangular(r2["up"], r2["dn"])
    got, ref = G.davidson_sparse_single(), oracle.davidson_single(cnt, idx, val)
    assert got["n_iter"] == ref["n_iter"] and np.max(np.abs(got["ritz"] - ref["ritz"])) < 1e-8
    assert abs(got["highest_eigenvalue"] - ref["highest"]) < 1e-8
