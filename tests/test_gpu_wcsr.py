"""The window-staged layout (csrc/wcsr.cu) against the oracle and against the plain CSR path.
SQMC_WCSR=1 forces the conversion on spaces the auto heuristic would leave in CSR (small / sparse
spaces: tiny tiles, many windows), which exercises the general code paths."""
import os

import numpy as np
import pytest

from conftest import C2_FCIDUMP, C2_ORBSYM

pytestmark = pytest.mark.gpu


@pytest.fixture()
def force_wcsr():
    old = os.environ.get("SQMC_WCSR")
    os.environ["SQMC_WCSR"] = "1"
    yield
    if old is None:
        del os.environ["SQMC_WCSR"]
    else:
        os.environ["SQMC_WCSR"] = old


def _check_against_oracle(oracle, H, ref, tol=1e-12):
    cnt, idx, val = ref
    n = len(cnt)
    got = H.export_upper()  # host decode of the window-major layout
    assert np.array_equal(got[0], cnt) and np.array_equal(got[1], idx) and np.array_equal(got[2], val)
    rng = np.random.default_rng(99)
    x = rng.uniform(-1, 1, n)
    y, yref = H.matvec(x), oracle.matvec_upper(cnt, idx, val, x)
    assert np.max(np.abs(y - yref)) <= tol * np.max(np.abs(yref))
    ptr = np.zeros(n + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(cnt)
    A = oracle.upper_to_scipy(cnt, idx, val)
    for i in (0, n // 2, n - 1):
        gc, gv = H.get_row(i + 1)
        row = A.getrow(i)
        order = np.argsort(row.indices)
        assert np.array_equal(gc - 1, row.indices[order]) and np.array_equal(gv, row.data[order])


def test_wcsr_forced_on_sparse_c2_space(oracle, c2_space, force_wcsr):
    import sqmc_b200 as sq
    s, r = c2_space
    ref = s.build_upper(r["up"], r["dn"])
    H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP))
    assert H.generate_sparse_ham_upper_triangular(r["up"], r["dn"]) == len(ref[1])
    _check_against_oracle(oracle, H, ref)
    d = H.davidson_sparse(n_states=1)
    dref = oracle.davidson(*ref, n_states=1)
    assert np.max(np.abs(d["ritz"] - dref["ritz"])) < 1e-8
    tau, e_trial = 0.01, float(ref[2][0])
    H.scale_values(-tau)
    w = np.ones(len(ref[0])) / np.sqrt(len(ref[0]))
    dw = H.projector_step(tau, e_trial, w)
    _, dwr = oracle.projector_step(ref[0], ref[1], -tau * ref[2], tau, e_trial, w)
    assert np.max(np.abs(dw - dwr)) <= 1e-12 * np.max(np.abs(dwr)) + 1e-15


def test_wcsr_forced_on_heg_and_hubbard(oracle, heg_space, force_wcsr):
    import itertools
    import sqmc_b200 as sq
    s, r = heg_space
    ref = s.build_upper(r["up"], r["dn"])
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49))
    H.generate_sparse_ham_upper_triangular(r["up"], r["dn"])
    _check_against_oracle(oracle, H, ref)
    hs = sq.HubbardKSystem(4, 4, 1.0, 4.0, 3, 3)
    so = oracle.System.hubbardk(4, 4, 1.0, 4.0, 3, 3)
    strings = [sum(1 << o for o in c) for c in itertools.combinations(range(16), 3)]
    dets = sorted((u, d) for u in strings for d in strings if hs.total_momentum(u, d) == (0, 0))
    up = oracle.dets_to_u64([u for u, d in dets])
    dn = oracle.dets_to_u64([d for u, d in dets])
    ref = so.build_upper(up, dn)
    H2 = sq.SparseHamiltonian(hs)
    H2.generate_sparse_ham_upper_triangular(up, dn)
    _check_against_oracle(oracle, H2, ref)


def test_wcsr_dense_space_matches_csr_path(oracle):
    """300k lowest-energy C2 determinants: dense enough for the automatic conversion (>= 64 dets per string);
    WCSR and CSR paths must agree, and both with the oracle on a sample of rows."""
    import sqmc_b200 as sq
    from sqmc_b200 import spaces
    chem = sq.ChemSystem(C2_FCIDUMP)
    up, dn, _ = spaces.c2_lowest_energy_space(chem, 300000)
    n = len(up)
    x = spaces.splitmix_vector(n, 7)
    os.environ["SQMC_WCSR"] = "0"
    H0 = sq.SparseHamiltonian(chem)
    nnz0 = H0.generate_sparse_ham_upper_triangular(up, dn)
    y0 = H0.matvec(x)
    os.environ["SQMC_WCSR"] = "1"
    H1 = sq.SparseHamiltonian(chem)
    nnz1 = H1.generate_sparse_ham_upper_triangular(up, dn)
    y1 = H1.matvec(x)
    del os.environ["SQMC_WCSR"]
    assert nnz0 == nnz1
    assert np.max(np.abs(y0 - y1)) <= 1e-13 * np.max(np.abs(y0))
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM)
    for i in (0, 12345, n - 1):
        rc, rv = S.row(up, dn, i)
        for H in (H0, H1):
            gc, gv = H.get_row(i + 1)
            assert np.array_equal(gc, rc) and np.array_equal(gv, rv)
    d0, d1 = H0.davidson_sparse(n_states=1), H1.davidson_sparse(n_states=1)
    assert abs(d0["evals"][0] - d1["evals"][0]) < 1e-9
