"""Incremental H build (chemistry.f90:7769-7843 `sparse_ham%ndet`): build_h with ndet_old > 0 keeps
the previous matrix and generates only pairs that involve a new determinant.  At every iteration of an HCI loop the
exported matrix must be bit-identical to the oracle's from-scratch build of the same list."""
import numpy as np
import pytest

import sqmc_b200 as sq
from conftest import C2_FCIDUMP

pytestmark = pytest.mark.gpu


def _check(H, S, up, dn, n_prev, want_incremental):
    nnz = H.generate_sparse_ham_upper_triangular(up, dn, ndet_old=n_prev)
    assert H.last_build_incremental() == want_incremental
    cnt, idx, val = S.build_upper(up, dn)
    assert nnz == len(idx)
    g = H.export_upper()
    assert np.array_equal(g[0], cnt) and np.array_equal(g[1], idx) and np.array_equal(g[2], val)


def test_c2_hci_loop_incremental_equals_from_scratch(oracle, c2_space):
    S, r = c2_space
    up, dn, sizes = r["up"], r["dn"], [int(v) for v in r["ndet"]]
    H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP, time_sym=False, z=1))
    prev = 0
    for n in sizes:
        _check(H, S, up[:n], dn[:n], prev, want_incremental=prev > 0)
        prev = n
    # H.v and Davidson on the extended matrix
    cnt, idx, val = S.build_upper(up, dn)
    x = np.random.default_rng(3).uniform(-1, 1, len(cnt))
    yref = oracle.matvec_upper(cnt, idx, val, x)
    assert np.max(np.abs(H.matvec(x) - yref)) <= 1e-12 * np.max(np.abs(yref))
    assert abs(H.davidson_sparse(n_states=1)["evals"][0] - oracle.davidson(cnt, idx, val, n_states=1)["evals"][0]) < 1e-8
    H.close()


def test_heg_hci_loop_incremental_equals_from_scratch(oracle, heg_space):
    S, r = heg_space
    up, dn, sizes = r["up"], r["dn"], [int(v) for v in r["ndet"]]
    assert sizes == [277, 9475]                          # the reference's golden counts (o_det_ref:261,330)
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49))
    prev = 0
    for n in [1] + sizes:
        _check(H, S, up[:n], dn[:n], prev, want_incremental=prev > 0)
        prev = n
    H.close()


def test_many_small_extensions_and_fallbacks(oracle, c2_space, c2_space_ts):
    """ragged growth (one determinant, a few, many; new determinants landing before / between / after the old ones in the
    internal order) and the conditions under which the library must rebuild instead"""
    S, r = c2_space
    up, dn = r["up"], r["dn"]
    n_all = len(up)
    H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP, time_sym=False, z=1))
    prev = 0
    for n in (1, 2, 3, 40, 41, 1000, 1001, 1500, 4000, n_all - 1, n_all):
        _check(H, S, up[:n], dn[:n], prev, want_incremental=prev > 0)
        prev = n
    # a different storage order in between does not matter (the library decodes before it merges)
    H.generate_sparse_ham_upper_triangular(up[:2000], dn[:2000])
    H.set_row_bundle(2)
    _check(H, S, up[:5000], dn[:5000], 2000, want_incremental=True)
    # scaled values (the walk keeps -tau*H): the previous matrix is not H any more -> rebuild
    H.scale_values(-0.01)
    _check(H, S, up[:6000], dn[:6000], 5000, want_incremental=False)
    # ndet_old that is not the previous list length -> rebuild
    _check(H, S, up[:7000], dn[:7000], 5999, want_incremental=False)
    H.close()
    # time-reversal symmetrised determinants: rebuilt (the entry list carries the time-reversed partners)
    St, rt = c2_space_ts
    Ht = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP, time_sym=True, z=1))
    sizes = [int(v) for v in rt["ndet"]]
    _check(Ht, St, rt["up"][:sizes[0]], rt["dn"][:sizes[0]], 0, want_incremental=False)
    _check(Ht, St, rt["up"][:sizes[1]], rt["dn"][:sizes[1]], sizes[0], want_incremental=False)
    Ht.close()


def test_two_word_strings_incremental(oracle):
    """81 plane waves (two 64-bit words per string): extension in three steps, each compared with the oracle from scratch"""
    S = oracle.System.heg(3, 0.5, 14, 7, 2.5)
    r = S.hci(2e-3, n_states=1, max_iters=1)
    up, dn = r["up"], r["dn"]
    n = len(up)
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 2.5))
    prev = 0
    for m in (n // 3, n // 2, n):
        _check(H, S, up[:m], dn[:m], prev, want_incremental=prev > 0)
        prev = m
    H.close()
