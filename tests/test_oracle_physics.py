"""Independent physics checks of the chem restatement (the reference ships no output for C2_v2z_curve):
an unrelated second-quantisation implementation (operators applied to occupation vectors, straight
from the FCIDUMP integrals) must give the same spectrum on small complete spaces, and the time-reversal
symmetrised basis must give the same energies as the plain determinant basis."""
import itertools
import json
import os

import numpy as np
import pytest

from conftest import C2_FCIDUMP, C2_ORBSYM

HERE = os.path.dirname(os.path.abspath(__file__))


def _second_quantised_h(chem, dets, nact):
    """H_ij = <i|H|j> by applying H = sum h_pq a+_p a_q + 1/2 sum (pq|rs) a+_p a+_r a_s a_q to |j>,
    spin orbitals ordered (all up, then all dn) -- an ordering convention of its own; only
    eigenvalues are compared."""
    n1 = chem.norb + 1
    h1 = np.array([[chem.integral(p + 1, q + 1, n1, n1) for q in range(nact)] for p in range(nact)])
    eri = np.zeros((nact,) * 4)
    for p, q, r, s in itertools.product(range(nact), repeat=4):
        eri[p, q, r, s] = chem.integral(p + 1, q + 1, r + 1, s + 1)
    index = {d: k for k, d in enumerate(dets)}

    def apply(ops, occ):  # ops: list of (spin_orbital, create?) applied right to left
        occ = list(occ)
        sign = 1
        for so, cr in reversed(ops):
            if cr == occ[so]:
                return 0, None
            sign *= -1 if sum(occ[:so]) % 2 else 1
            occ[so] = 1 if cr else 0
        return sign, tuple(occ)

    def occ_of(d):
        up, dn = d
        return tuple([(up >> k) & 1 for k in range(nact)] + [(dn >> k) & 1 for k in range(nact)])

    def det_of(occ):
        up = sum(occ[k] << k for k in range(nact))
        dn = sum(occ[nact + k] << k for k in range(nact))
        return up, dn

    H = np.zeros((len(dets), len(dets)))
    for j, dj in enumerate(dets):
        oj = occ_of(dj)
        for sp in (0, 1):
            for p in range(nact):
                for q in range(nact):
                    if h1[p, q] == 0.0:
                        continue
                    sg, o = apply([(sp * nact + p, True), (sp * nact + q, False)], oj)
                    if sg and det_of(o) in index:
                        H[index[det_of(o)], j] += sg * h1[p, q]
        for s1 in (0, 1):
            for s2 in (0, 1):
                for p, q, r, s in itertools.product(range(nact), repeat=4):
                    v = eri[p, q, r, s]
                    if v == 0.0:
                        continue
                    sg, o = apply([(s1 * nact + p, True), (s2 * nact + r, True), (s2 * nact + s, False), (s1 * nact + q, False)], oj)
                    if sg and det_of(o) in index:
                        H[index[det_of(o)], j] += 0.5 * sg * v
    return H + chem.enuc * np.eye(len(dets))


def test_chem_elements_against_second_quantisation(oracle):
    import sqmc_b200 as sq
    chem = sq.ChemSystem(C2_FCIDUMP)
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM)
    nact = 6  # complete space of 8 electrons in the 6 lowest (reordered) orbitals: 15 x 15 determinants
    strs = [sum(1 << o for o in c) for c in itertools.combinations(range(nact), 4)]
    dets = sorted((u, d) for u in strs for d in strs)
    up = oracle.dets_to_u64([u for u, d in dets])
    dn = oracle.dets_to_u64([d for u, d in dets])
    cnt, idx, val = S.build_upper(up, dn)
    A = oracle.upper_to_scipy(cnt, idx, val).toarray()
    B = _second_quantised_h(chem, dets, nact)
    assert np.allclose(B, B.T, atol=1e-12)
    assert np.allclose(np.abs(A), np.abs(B), atol=1e-10)          # element magnitudes
    assert np.allclose(np.linalg.eigvalsh(A), np.linalg.eigvalsh(B), atol=1e-9)  # phases consistent up to a basis sign change
    # HF energy = <HF|H|HF> from the FCIDUMP
    assert abs(A[0, 0] - B[0, 0]) < 1e-10


def test_time_sym_basis_gives_same_singlet_energy(oracle):
    """Ground state of C2 in the full 8e/6o space is a singlet (z=+1): the symmetrised basis
    (chemistry.f90:1323-1377) must reproduce the determinant-basis eigenvalue."""
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM)
    T = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=True, z=1)
    nact = 7
    strs = [sum(1 << o for o in c) for c in itertools.combinations(range(nact), 4)]
    dets = sorted((u, d) for u in strs for d in strs)
    reps = [(u, d) for (u, d) in dets if u <= d]
    A = oracle.upper_to_scipy(*S.build_upper(oracle.dets_to_u64([u for u, d in dets]), oracle.dets_to_u64([d for u, d in dets]))).toarray()
    B = oracle.upper_to_scipy(*T.build_upper(oracle.dets_to_u64([u for u, d in reps]), oracle.dets_to_u64([d for u, d in reps]))).toarray()
    assert np.allclose(B, B.T, atol=1e-13)
    wa, wb = np.linalg.eigvalsh(A), np.linalg.eigvalsh(B)
    assert abs(wa[0] - wb[0]) < 1e-9
    assert all(np.min(np.abs(wa - e)) < 1e-9 for e in wb)  # every z=+1 eigenvalue is an eigenvalue of H


def test_c2_hci_fixture_is_reproducible(oracle, c2_space_ts):
    """tests/golden/c2_s1_hci.json (oracle-generated, committed) still agrees with the oracle."""
    gold = json.load(open(os.path.join(HERE, "golden", "c2_s1_hci.json")))["runs"]["n_states=1"]
    s, r = c2_space_ts  # first 3 iterations of the same run
    assert r["ndet"].tolist() == gold["n_det"][:3]
    assert r["nnz"].tolist() == gold["nnz"][:3]
    assert np.max(np.abs(r["iter_energy"][:, 0] - np.array(gold["iter_energy"])[:3, 0])) < 1e-9
