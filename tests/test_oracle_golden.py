"""Pins the CPU oracle on the reference's OWN golden log (src/e2e_tests/heg/o_det_ref),
transcribed to tests/golden/heg_o_det_ref.json by tests/golden/make_golden.py."""
import json
import os

import numpy as np
import pytest

from conftest import label_sorted

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "heg_o_det_ref.json")))


def test_heg_orbital_order_matches_reference_log(oracle):
    s = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    assert s.norb == 19  # o_det_ref:52
    kv = s.heg_tables()["k_vectors"]
    for k, row in enumerate(GOLD["k_points"]):  # o_det_ref:53-71 (printed with 4 decimals)
        assert np.allclose(kv[k], row[:3], atol=5.1e-5), (k, kv[k], row)
        assert abs(np.sqrt((kv[k] ** 2).sum()) - row[3]) < 5.1e-6


def test_heg_hf_energy(oracle):
    s = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    hf = oracle.dets_to_u64([127])
    assert abs(s.elements(hf, hf, hf, hf)[0] - GOLD["hf_energy"]) < 5.1e-9  # o_det_ref:74


def test_heg_hci_reproduces_reference_log(oracle, heg_space):
    s, r = heg_space
    assert r["ndet"].tolist() == GOLD["n_det"]      # 277, 9475   (o_det_ref:261,330)
    assert r["nnz"].tolist() == GOLD["nnz"]         # 3511, 165193
    for it in range(2):                             # per-iteration Davidson Ritz values (o_det_ref:270-274,339-344)
        got = r["ritz"][it].ravel()
        ref = np.array(GOLD["ritz"][it])
        assert len(got) == len(ref)
        assert np.max(np.abs(got - ref)) < 1.1e-9   # printed with 9 decimals
        assert abs(r["iter_energy"][it, 0] - GOLD["davidson_final"][it]["energy"]) < 1.1e-10  # o_det_ref:275,345
    # CI coefficients incl. signs (o_det_ref:394-413): pins the fermionic phases
    ups = oracle.u64_to_ints(r["up"])
    dns = oracle.u64_to_ints(r["dn"])
    where = {(u, d): k for k, (u, d) in enumerate(zip(ups, dns))}
    for c in GOLD["final_coefficients"]:
        k = where[(c["up"], c["dn"])]
        assert abs(r["wts"][k, 0] - c["coef"]) < 1e-12


def test_incremental_build_equals_full_build(oracle, heg_space):
    """sparse_ham reuse (chemistry.f90:7818-7843 / heg.f90:3724): rows 1..ndet_old copied, new links appended."""
    s, r = heg_space
    up, dn = r["up"], r["dn"]
    s2 = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    a = s2.build_upper(up[:277], dn[:277])
    assert len(a[1]) == 3511
    b = s2.build_upper(up, dn, incremental=True)
    s3 = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    c = s3.build_upper(up, dn)
    for x, y in zip(b, c):
        assert np.array_equal(x, y)


def test_davidson_against_dense(oracle, heg_space):
    s, r = heg_space
    s2 = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    cnt, idx, val = s2.build_upper(r["up"][:277], r["dn"][:277])
    A = oracle.upper_to_scipy(cnt, idx, val).toarray()
    assert np.allclose(A, A.T)
    w = np.linalg.eigvalsh(A)
    d = oracle.davidson(cnt, idx, val, n_states=1)
    assert abs(d["evals"][0] - w[0]) < 1e-9
    x = np.random.default_rng(3).uniform(-1, 1, 277)
    assert np.allclose(oracle.matvec_upper(cnt, idx, val, x), A @ x, rtol=0, atol=1e-12)
    assert np.allclose(oracle.matvec_upper_mt(cnt, idx, val, x, 3), A @ x, rtol=0, atol=1e-12)


def test_davidson_two_states_against_dense(oracle):
    """n_states=2 on the committed C2 fixture (2000 lowest-energy A_g determinants).
    NOTE: the reference's multi-state Davidson appends the (normalised) residual of an already
    converged state; on matrices where one state converges long before the other (e.g. the 277-det
    HEG space started from unit vectors) the basis loses orthogonality and the Ritz values collapse.
    That is reference behaviour (more_tools.f90:2165-2184) which the restatement keeps."""
    g = np.load(os.path.join(HERE, "golden", "c2_small_space.npz"))
    cnt, idx, val = g["counts"], g["indices"], g["values"]
    w = np.linalg.eigvalsh(oracle.upper_to_scipy(cnt, idx, val).toarray())
    d = oracle.davidson(cnt, idx, val, n_states=2)
    assert np.max(np.abs(d["evals"] - w[:2])) < 1e-9


def test_lanczos_against_dense(oracle, heg_space):
    """matrix_lanczos_sparse restatement (more_tools.f90:1742-1883): lowest eigenpair against numpy on the 277-det matrix,
    whose Davidson energy the reference log pins (o_det_ref:270)."""
    s, r = heg_space
    up, dn = r["up"][:277], r["dn"][:277]
    cnt, idx, val = s.build_upper(up, dn)
    A = oracle.upper_to_scipy(cnt, idx, val).toarray()
    w, v = np.linalg.eigh(A)
    L = oracle.lanczos(cnt, idx, val)
    assert abs(L["lowest"] - w[0]) < 1e-9 and abs(L["lowest"] - 58.2825967049) < 5e-9
    assert L["n_iter"] == len(L["ritz"]) + 1                      # the converging step is not printed (:1847-1853)
    assert abs(abs(np.dot(L["evec"], v[:, 0])) - 1) < 1e-6
    x = np.sin(np.arange(277.0)) + 2.0
    L2 = oracle.lanczos(cnt, idx, val, v0=x)
    assert abs(L2["lowest"] - w[0]) < 1e-9 and L2["highest"] <= w[-1] + 1e-9 and L2["second_lowest"] >= w[1] - 1e-9


def test_heg_pt_reproduces_reference_log(oracle):
    """second_order_pt restatement against the reference's own log (o_det_ref:431,437): eps_pt = 2e-7 on the final
    9475-determinant wavefunction -> 501881 connected determinants, PT correction -0.000939196, total 58.275966889."""
    gold = json.load(open(os.path.join(HERE, "golden", "heg_o_det_ref.json")))["pt"]
    S = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    r = S.hci(1e-3, n_states=1)
    assert len(r["up"]) == gold["ndets"]
    de, nconn = S.pt2(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], gold["eps_pt"])
    assert nconn == gold["ndets_connected"]
    assert abs(de - gold["pt_correction"]) < 5e-10                  # printed with 9 decimals
    assert abs(r["energy"][0] + de - gold["total_energy"]) < 1e-9
    # the semistochastic-PT test of the same system pins the deterministic stage at a second threshold (o_st_ref:432)
    big = json.load(open(os.path.join(HERE, "golden", "heg_o_det_ref.json")))["pt_big"]
    de2, nconn2 = S.pt2(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], big["eps_pt_big"])
    assert nconn2 == big["ndets_connected"] == 13159 and abs(de2 - big["pt_correction"]) < 5e-10


def test_heg_stochastic_pt_reproduces_reference_log(oracle):
    """second_order_pt_alias restatement (hci.f90:1314-1684: rannyu stream, alias tables, n_mc = 200 draws per sample,
    term1/term2 with the eps_pt_big parts removed, Welford) against the reference's own log src/e2e_tests/heg/o_st_ref:
    all 143 printed samples (E_2pt_now to the printed 9 decimals, number of distinct sampled determinants), the stopping
    sample, the final estimate -0.000729402 +- 0.000009966 and the total PT lowering -0.000928741 (o_st_ref:442-875)."""
    gold = json.load(open(os.path.join(HERE, "golden", "heg_o_det_ref.json")))
    g = gold["pt_stochastic"]
    S = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    r = S.hci(1e-3, n_states=1)
    up, dn, w = label_sorted(r)
    res = S.pt2_alias(up, dn, w, r["energy"][0], g["eps_pt"], g["eps_pt_big"], g["n_mc"], g["target_error"], g["irand_seed_1"], max_samples=400)
    assert len(res["e_now"]) == len(g["samples"]) == 143                      # same stopping sample
    assert res["n_distinct"].tolist() == g["n_ref"]                            # same draws
    assert int(res["n_connected"][-1]) == g["ndets_connected_last_sample"] == 23726   # o_st_ref:882 "ndets_connected(total)"
    assert np.max(np.abs(res["e_now"] - np.array([x["e_now"] for x in g["samples"]]))) < 5.1e-10
    assert abs(res["pt_energy"] - g["pt_diff"]) < 5.1e-10 and abs(res["std_dev"] - g["std_dev"]) < 5.1e-10
    de_big, _ = S.pt2(up, dn, w, r["energy"][0], g["eps_pt_big"])
    assert abs(de_big + res["pt_energy"] - g["pt_total"]) < 1.1e-9


def test_davidson_single_against_dense(oracle, heg_space):
    """davidson_sparse_single restatement (more_tools.f90:3055-3233) on the 277-determinant matrix of the reference log:
    its printed eigenvalues coincide with davidson_sparse's golden ones (the solvers differ only in guard and restart)"""
    gold = json.load(open(os.path.join(HERE, "golden", "heg_o_det_ref.json")))
    s, r = heg_space
    up, dn = r["up"][:277], r["dn"][:277]
    cnt, idx, val = s.build_upper(up, dn)
    A = oracle.upper_to_scipy(cnt, idx, val).toarray()
    w = np.linalg.eigvalsh(A)
    L = oracle.davidson_single(cnt, idx, val)
    assert abs(L["lowest"] - w[0]) < 1e-9
    assert np.max(np.abs(L["ritz"] - np.array(gold["ritz"][0][:len(L["ritz"])]))) < 5e-9      # o_det_ref, first HCI iteration
    assert abs(L["highest"] - max(A.diagonal().max(), L["highest"])) == 0 and L["highest"] <= w[-1] + 1e-9
