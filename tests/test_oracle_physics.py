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


@pytest.mark.parametrize("time_sym", [False, True])
def test_pt2_numerators_equal_full_matrix_elements(oracle, time_sym):
    """Independent check of the second-order PT restatement (parity of C2 PT is unpinned by the reference): with eps_pt -> 0
    the screened sum must equal the textbook Epstein-Nesbet sum over ALL singles and doubles of the variational
    determinants, with numerators built from the full (for time_sym: symmetrised) matrix elements -- i.e. the
    contribution-by-contribution bookkeeping of find_important_connected_dets_chem (norm factors, z, merging of a
    determinant with its time-reversed partner) adds up to <a|H|psi>."""
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=time_sym, z=1, hf_symmetry=1)
    r = S.hci(5e-2, n_states=1, max_iters=1)
    up, dn, w, e = r["up"], r["dn"], r["wts"][:, 0], r["energy"][0]
    n = len(up)
    assert 10 < n < 200
    de, nconn = S.pt2(up, dn, w, e, 1e-13)
    V = {(int(u[0]), int(d[0])) for u, d in zip(up, dn)}
    norb = 26

    def excitations(u, d):
        occu = [o for o in range(norb) if u >> o & 1]; viru = [o for o in range(norb) if not u >> o & 1]
        occd = [o for o in range(norb) if d >> o & 1]; vird = [o for o in range(norb) if not d >> o & 1]
        out = set()
        for p in occu:
            for q in viru:
                out.add((u ^ (1 << p) | (1 << q), d))
        for p in occd:
            for q in vird:
                out.add((u, d ^ (1 << p) | (1 << q)))
        for p, q in itertools.combinations(occu, 2):
            for a, b in itertools.combinations(viru, 2):
                out.add((u ^ (1 << p) ^ (1 << q) | (1 << a) | (1 << b), d))
        for p, q in itertools.combinations(occd, 2):
            for a, b in itertools.combinations(vird, 2):
                out.add((u, d ^ (1 << p) ^ (1 << q) | (1 << a) | (1 << b)))
        for p in occu:
            for a in viru:
                nu = u ^ (1 << p) | (1 << a)
                for q in occd:
                    for b in vird:
                        out.add((nu, d ^ (1 << q) | (1 << b)))
        return out

    ext = set()
    for u, d in V:
        for a, b in excitations(u, d):
            if time_sym and a > b:
                a, b = b, a
            if (a, b) not in V:
                ext.add((a, b))
    ext = sorted(ext)
    m = len(ext)
    au = oracle.dets_to_u64([a for a, b in ext])
    ad = oracle.dets_to_u64([b for a, b in ext])
    num = np.zeros(m)
    for i in range(n):
        num += S.elements(au, ad, np.repeat(up[i:i + 1], m, axis=0), np.repeat(dn[i:i + 1], m, axis=0)) * w[i]
    haa = S.elements(au, ad, au, ad)
    ref = float(np.sum(num ** 2 / (e - haa)))
    assert abs(de - ref) < 1e-12 and de < -1e-2


def _all_excitations(u, d, norb):
    occu = [o for o in range(norb) if u >> o & 1]; viru = [o for o in range(norb) if not u >> o & 1]
    occd = [o for o in range(norb) if d >> o & 1]; vird = [o for o in range(norb) if not d >> o & 1]
    out = set()
    for p in occu:
        for q in viru:
            out.add((u ^ (1 << p) | (1 << q), d))
    for p in occd:
        for q in vird:
            out.add((u, d ^ (1 << p) | (1 << q)))
    for p, q in itertools.combinations(occu, 2):
        for a, b in itertools.combinations(viru, 2):
            out.add((u ^ (1 << p) ^ (1 << q) | (1 << a) | (1 << b), d))
    for p, q in itertools.combinations(occd, 2):
        for a, b in itertools.combinations(vird, 2):
            out.add((u, d ^ (1 << p) ^ (1 << q) | (1 << a) | (1 << b)))
    for p in occu:
        for a in viru:
            nu = u ^ (1 << p) | (1 << a)
            for q in occd:
                for b in vird:
                    out.add((nu, d ^ (1 << q) | (1 << b)))
    return out


def test_stochastic_pt_sample_equals_direct_formula(oracle):
    """Independent check of the stochastic-PT sample restatement on C2 (parity-unpinned by the reference): with eps_pt -> 0 and no
    'big' part (eps_pt_big huge) one sample must equal, term by term, the estimator of hci.f90:1616-1632 written directly with full
    matrix elements: sum_k [ (sum_i H_ki c_i w_i)^2 + sum_i (H_ki c_i)^2 ((n_mc-1) w_i - w_i^2) ] / (E - H_kk) / (n_mc (n_mc-1)),
    k over everything outside the variational list that the sampled determinants connect to."""
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=False, z=1, hf_symmetry=1)
    r = S.hci(5e-2, n_states=1, max_iters=1)
    up, dn, w, e = r["up"], r["dn"], r["wts"][:, 0], r["energy"][0]
    key = [(int(u[0]), int(d[0])) for u, d in zip(up, dn)]
    o = np.array(sorted(range(len(key)), key=lambda i: key[i]))
    up, dn, w = up[o], dn[o], w[o]
    n, n_mc = len(up), 7
    V = {(int(u[0]), int(d[0])) for u, d in zip(up, dn)}
    rng = np.random.default_rng(11)
    prob = np.abs(w) / np.abs(w).sum()
    idx, counts = np.unique(rng.choice(n, size=n_mc, p=prob), return_counts=True)
    wop = counts / prob[idx]
    est, nconn = S.pt2_sample(up, dn, up[idx], dn[idx], w[idx], wop, n_mc, e, 1e-13, 1e9)
    ext = sorted({x for i in idx for x in _all_excitations(int(up[i][0]), int(dn[i][0]), 26)} - V)
    m = len(ext)
    au = oracle.dets_to_u64([a for a, b in ext]); ad = oracle.dets_to_u64([b for a, b in ext])
    t1, t2 = np.zeros(m), np.zeros(m)
    for q, i in enumerate(idx):
        hc = S.elements(au, ad, np.repeat(up[i:i + 1], m, axis=0), np.repeat(dn[i:i + 1], m, axis=0)) * w[i]
        t1 += hc * wop[q]
        t2 += hc ** 2 * ((n_mc - 1) * wop[q] - wop[q] ** 2)
    haa = S.elements(au, ad, au, ad)
    ref = float(np.sum((t1 ** 2 + t2) / (e - haa))) / (n_mc * (n_mc - 1))
    assert abs(est - ref) < 1e-12 * max(1.0, abs(ref)) and est != 0.0


def test_stochastic_pt_is_unbiased(oracle):
    """Size-independent property of second_order_pt_alias: the mean of the per-sample energies estimates the DIFFERENCE between the
    deterministic corrections at eps_pt and eps_pt_big (hci.f90:1428: 'This is the difference between the PT correction for eps_pt
    and eps_pt_big').  600 samples of 40 determinants on a small C2 wavefunction: the mean must agree with pt2(eps_pt) - pt2(eps_pt_big)
    within 4 standard errors, and the standard error the loop reports must match the scatter of the samples."""
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=False, z=1, hf_symmetry=1)
    r = S.hci(2e-2, n_states=1, max_iters=1)
    up, dn, w, e = r["up"], r["dn"], r["wts"][:, 0], r["energy"][0]
    key = [(int(u[0]), int(d[0])) for u, d in zip(up, dn)]
    o = np.array(sorted(range(len(key)), key=lambda i: key[i]))
    up, dn, w = up[o], dn[o], w[o]
    eps_pt, eps_big = 1e-4, 2e-3
    exact = S.pt2(up, dn, w, e, eps_pt)[0] - S.pt2(up, dn, w, e, eps_big)[0]
    res = S.pt2_alias(up, dn, w, e, eps_pt, eps_big, 40, 0.0, [12, 34, 56, 79], max_samples=600)
    assert len(res["e_now"]) == 600                       # target_error = 0: runs to max_samples
    mean, sem = res["e_now"].mean(), res["e_now"].std(ddof=1) / np.sqrt(600)
    assert abs(res["pt_energy"] - mean) < 1e-15 and abs(res["std_dev"] - sem) < 1e-12 * max(1.0, sem)
    assert exact < 0 and abs(mean - exact) < 4 * sem, (mean, exact, sem)
