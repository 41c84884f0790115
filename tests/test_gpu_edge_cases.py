"""Edge cases: two-word determinant strings (65..127 orbitals, the reference's integer(16) limit, types.f90:44),
tiny spaces (n = 1, 2), and the C ABI's error behaviour."""
import os

import numpy as np
import pytest

from conftest import C2_FCIDUMP

pytestmark = pytest.mark.gpu


def _same(got, ref):
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


@pytest.mark.parametrize("cutoff,eps,norb", [(2.5, 2e-3, 81), (3.0, 3e-3, 123)])
def test_heg_two_word_strings(oracle, cutoff, eps, norb):
    """HEG with 81 and 123 plane-wave orbitals: strings need both 64-bit words (NW = 2 code paths everywhere)."""
    import sqmc_b200 as sq
    S = oracle.System.heg(3, 0.5, 14, 7, cutoff)
    assert S.norb == norb
    r = S.hci(eps, n_states=1, max_iters=1)
    up, dn = r["up"], r["dn"]
    ref = S.build_upper(up, dn)
    hs = sq.HegSystem(3, 0.5, 14, 7, cutoff)
    H = sq.SparseHamiltonian(hs)
    assert H.generate_sparse_ham_upper_triangular(up, dn) == len(ref[1])
    _same(H.export_upper(), ref)
    x = np.random.default_rng(3).uniform(-1, 1, len(up))
    y, yref = H.matvec(x), oracle.matvec_upper(*ref, x)
    assert np.max(np.abs(y - yref)) <= 1e-12 * np.max(np.abs(yref))
    d = H.davidson_sparse(n_states=1)
    assert abs(d["evals"][0] - r["iter_energy"][0, 0]) < 1e-8
    # selection step with two-word strings
    coeffs = np.abs(r["wts"][:, 0])
    mh = np.full(len(up), 9e99)
    ru, rd_, rm = S.select(up, dn, coeffs, mh, 0.3 * eps)
    gu, gd, gm = H.get_next_det_list(up, dn, coeffs, mh, 0.3 * eps)
    assert len(ru) > 0 and np.array_equal(gu, ru) and np.array_equal(gd, rd_) and np.array_equal(gm, rm)


def test_chem_two_word_strings_synthetic(oracle, tmp_path):
    """A synthetic 70-orbital FCIDUMP (random symmetric integrals, no physics): exercises the chem element code with
    two-word strings, single / double / time-reversal branches, against the oracle."""
    import itertools
    import sqmc_b200 as sq
    norb, nelec = 70, 4
    rng = np.random.default_rng(2024)
    path = str(tmp_path / "FCIDUMP")
    with open(path, "w") as f:
        f.write(" &FCI NORB=%d,NELEC=%d,MS2=0,\n  ORBSYM=%s\n  ISYM=1,\n &END\n" % (norb, nelec, "1," * norb))
        orbs = list(range(1, norb + 1))
        for _ in range(6000):
            p, q, r, s = (int(v) for v in rng.choice(orbs, 4))
            f.write("%.12f %d %d %d %d\n" % (rng.uniform(-0.5, 0.5), p, q, r, s))
        for p in orbs:
            f.write("%.12f %d %d 0 0\n" % (rng.uniform(-2, -0.1) + 0.05 * p, p, p))
            if p < norb:
                f.write("%.12f %d %d 0 0\n" % (rng.uniform(-0.2, 0.2), p, p + 1))
        f.write("1.5 0 0 0 0\n")
    for ts in (False, True):
        cs = sq.ChemSystem(path, time_sym=ts, z=1)
        S = oracle.System.chem(path, norb, nelec, 2, [1] * norb, time_sym=ts, z=1)
        assert np.array_equal(cs.integrals, S.chem_tables()["integrals"])
        # determinants: pairs of 2-electron strings touching the high orbitals (> 64)
        strs = [sum(1 << o for o in c) for c in itertools.combinations([0, 1, 2, 30, 63, 64, 65, 69], 2)]
        dets = sorted((u, d) for u in strs for d in strs if (not ts or u <= d))
        up = oracle.dets_to_u64([u for u, d in dets])
        dn = oracle.dets_to_u64([d for u, d in dets])
        ref = S.build_upper(up, dn)
        H = sq.SparseHamiltonian(cs)
        assert H.generate_sparse_ham_upper_triangular(up, dn) == len(ref[1])
        _same(H.export_upper(), ref)
        assert np.array_equal(H.diagonal(up, dn), S.elements(up, dn, up, dn))


def test_tiny_spaces(oracle):
    import sqmc_b200 as sq
    cs = sq.ChemSystem(C2_FCIDUMP)
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, cs.orbital_symmetries_fcidump)
    H = sq.SparseHamiltonian(cs)
    hf = sq.dets_to_u64([cs.hf_up])
    assert H.generate_sparse_ham_upper_triangular(hf, hf) == 1          # n = 1: the diagonal only
    cnt, idx, val = H.export_upper()
    assert cnt.tolist() == [1] and idx.tolist() == [1] and val[0] == S.elements(hf, hf, hf, hf)[0]
    d = H.davidson_sparse(n_states=1)                                   # "Diagonalization attempted with n=1" (more_tools.f90:2235)
    assert d["evals"][0] == val[0]
    up = sq.dets_to_u64([cs.hf_up, cs.hf_up])
    dn = sq.dets_to_u64([cs.hf_dn, (cs.hf_dn & ~1) | (1 << 8)])         # a dn single excitation
    ref = S.build_upper(up, dn)
    assert H.generate_sparse_ham_upper_triangular(up, dn) == len(ref[1])
    _same(H.export_upper(), ref)
    x = np.array([0.3, -0.7])
    assert np.allclose(H.matvec(x), oracle.matvec_upper(*ref, x), rtol=0, atol=1e-13)


def test_error_behaviour():
    import sqmc_b200 as sq
    cs = sq.ChemSystem(C2_FCIDUMP)
    H = sq.SparseHamiltonian(cs)
    with pytest.raises(sq.SqmcError):                                   # no matrix yet
        H.n = 3
        H.matvec(np.zeros(3))
    bad = sq.dets_to_u64([1 << 40])                                     # orbital 41 > norb = 26
    with pytest.raises(sq.SqmcError) as e:
        H.generate_sparse_ham_upper_triangular(bad, bad)
    assert "beyond norb" in str(e.value)
    hf = sq.dets_to_u64([cs.hf_up])
    with pytest.raises(sq.SqmcError):                                   # ndet_old larger than n
        H.generate_sparse_ham_upper_triangular(hf, hf, ndet_old=5)
    hub = sq.HubbardKSystem(4, 4, 1.0, 4.0, 3, 3)
    with pytest.raises(sq.SqmcError):                                   # selection exists for chem / heg only (hci.f90:1073)
        sq.SparseHamiltonian(hub).get_next_det_list(hf, hf, [1.0], [9e99], 1e-3)


def test_alloc_stall_counter():
    """sqmc_b200_alloc_stall_ms: host time inside allocator / mapping calls is accumulated and can be reset"""
    import sqmc_b200 as sq
    from sqmc_b200 import _lib, spaces
    L = _lib.load()
    L.sqmc_b200_alloc_stall_ms(1)
    chem = sq.ChemSystem(C2_FCIDUMP)
    up, dn, _ = spaces.c2_lowest_energy_space(chem, 2000)
    H = sq.SparseHamiltonian(chem)
    H.generate_sparse_ham_upper_triangular(up, dn)
    v = L.sqmc_b200_alloc_stall_ms(1)
    assert v > 0.0 and v < 60000.0          # a build allocates
    assert L.sqmc_b200_alloc_stall_ms(0) == 0.0
    H.close()
