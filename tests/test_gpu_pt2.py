"""Second-order PT on the device (csrc/select.cu: pt2, pt2_sample, pt2_alias) against the reference's golden logs and the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import C2_FCIDUMP, label_sorted

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_heg_pt_reproduces_reference_log(oracle):
    """src/e2e_tests/heg/o_det_ref:431-437: eps_pt = 2e-7 on the 9475-determinant variational wavefunction:
    501881 connected determinants, PT correction -0.000939196, total energy 58.275966889"""
    import sqmc_b200 as sq
    S = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    r = S.hci(1e-3, n_states=1)
    assert len(r["up"]) == 9475
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49))
    de, nconn = H.second_order_pt(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], 2e-7)
    assert nconn == 501881
    assert abs(de - (-0.000939196)) < 5e-10
    assert abs(r["energy"][0] + de - 58.275966889) < 1e-9
    ode, onc = S.pt2(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], 2e-7)
    assert onc == nconn and abs(ode - de) < 1e-12
    # second golden threshold: the deterministic stage of the semistochastic-PT test (src/e2e_tests/heg/o_st_ref:432)
    de2, nconn2 = H.second_order_pt(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], 8.192e-4)
    assert nconn2 == 13159 and abs(de2 - (-0.000199339)) < 5e-10


def test_heg_pt_fully_on_gpu():
    """variational stage and PT both through the library (no oracle wavefunction): same golden numbers"""
    import sqmc_b200 as sq
    from sqmc_b200 import spaces
    hs = sq.HegSystem(3, 0.5, 14, 7, 1.49)
    H = sq.SparseHamiltonian(hs)
    up, dn, wts, e = spaces.hci_space(H, hs, 10**9, eps_schedule=[1e-3] * 4)
    assert len(up) == 9475 and abs(e - 58.276906085) < 5e-9           # o_det_ref:437
    de, nconn = H.second_order_pt(up, dn, wts[:, 0], e, 2e-7)
    assert nconn == 501881 and abs(e + de - 58.275966889) < 5e-9


@pytest.mark.parametrize("eps_pt", [1e-5, 1e-6])
def test_c2_pt_matches_oracle(oracle, c2_space, eps_pt):
    """C2 cc-pVDZ, plain determinants (time_sym = f): singles + table-driven doubles against the oracle's enumeration"""
    import sqmc_b200 as sq
    s, r = c2_space
    H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP, time_sym=False, z=1))
    de, nconn = H.second_order_pt(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], eps_pt)
    ode, onc = s.pt2(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], eps_pt)
    assert nconn == onc
    assert abs(de - ode) < 1e-11 and de < 0


@pytest.mark.parametrize("eps_pt", [1e-5, 1e-6])
def test_c2_time_sym_pt_matches_oracle(oracle, c2_space_ts, eps_pt):
    """the shipped C2 input is time_sym = t: numerators carry the norm factors and z of chemistry.f90:6961-6971,7121-7134,
    contributions of a determinant and of its time-reversed partner merge on the representative, H_aa is the
    symmetrised diagonal element (hci.f90:1164-1166)"""
    import sqmc_b200 as sq
    s, r = c2_space_ts
    H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP, time_sym=True, z=1))
    de, nconn = H.second_order_pt(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], eps_pt)
    ode, onc = s.pt2(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], eps_pt)
    assert nconn == onc
    assert abs(de - ode) < 1e-11 and de < 0


def test_pt_rejects_bad_arguments(c2_space):
    import sqmc_b200 as sq
    s, r = c2_space
    H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP))
    with pytest.raises(Exception, match="eps_pt"):
        H.second_order_pt(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], 0.0)


def test_heg_two_word_strings_pt(oracle):
    """81 plane-wave orbitals: determinant strings span both 64-bit words (NW = 2 instantiation of the PT pipeline)"""
    import sqmc_b200 as sq
    S = oracle.System.heg(3, 0.5, 14, 7, 2.5)
    r = S.hci(2e-3, n_states=1, max_iters=1)
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 2.5))
    de, nconn = H.second_order_pt(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], 2e-5)
    ode, onc = S.pt2(r["up"], r["dn"], r["wts"][:, 0], r["energy"][0], 2e-5)
    assert nconn == onc and onc > len(r["up"])
    assert abs(de - ode) < 1e-12 and de < 0


def test_heg_stochastic_pt_reproduces_reference_log(oracle):
    """second_order_pt_alias on the device against the reference's own log src/e2e_tests/heg/o_st_ref:442-875: seeded with the
    log's irand_seed, n_mc = 200, eps_pt = 2e-7, eps_pt_big = 8.192e-4, target_error = 1e-5 the call must print the same 143
    samples (E_2pt_now to the 9 printed decimals), stop at the same sample and end at -0.000729402 +- 0.000009966; with the
    deterministic eps_pt_big stage the PT lowering is -0.000928741 and the total 58.275977344."""
    import sqmc_b200 as sq
    g = json.load(open(os.path.join(HERE, "golden", "heg_o_det_ref.json")))["pt_stochastic"]
    S = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    r = S.hci(1e-3, n_states=1)
    up, dn, w = label_sorted(r)
    H = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49))
    seed = list(g["irand_seed_1"]); seed[3] = 2 * (seed[3] // 2) + 1          # setrn (rannyu.f90:19)
    res = H.second_order_pt_alias(up, dn, w, r["energy"][0], g["eps_pt"], g["eps_pt_big"], g["n_mc"], g["target_error"], seed, max_samples=400)
    gold = np.array([x["e_now"] for x in g["samples"]])
    assert res["n_samples"] == len(gold) == 143
    assert np.max(np.abs(res["e_2pt_samples"] - gold)) < 5.1e-10
    assert abs(res["pt_energy"] - g["pt_diff"]) < 5.1e-10 and abs(res["pt_energy_std_dev"] - g["std_dev"]) < 5.1e-10
    de_big, _ = H.second_order_pt(up, dn, w, r["energy"][0], g["eps_pt_big"])
    assert abs(de_big + res["pt_energy"] - g["pt_total"]) < 1.1e-9
    assert abs(r["energy"][0] + de_big + res["pt_energy"] - 58.275977344) < 2e-9
    # the oracle's restatement of the same loop: same samples to rounding, and the generator state moved on
    o = S.pt2_alias(up, dn, w, r["energy"][0], g["eps_pt"], g["eps_pt_big"], g["n_mc"], g["target_error"], seed, max_samples=400)
    assert np.max(np.abs(res["e_2pt_samples"] - o["e_now"])) < 1e-15
    assert list(res["rannyu_state"]) != seed and res["rannyu_state"][3] % 2 == 1


def _draw_sample(w, n_mc, seed):
    rng = np.random.default_rng(seed)
    prob = np.abs(w) / np.abs(w).sum()
    idx, counts = np.unique(rng.choice(len(w), size=n_mc, p=prob), return_counts=True)
    return idx, counts / prob[idx]


@pytest.mark.parametrize("ts", [False, True])
def test_c2_pt_sample_matches_oracle(oracle, c2_space, c2_space_ts, ts):
    """one stochastic-PT sample on C2 cc-pVDZ (plain and time-reversal symmetrised determinants): the four per-determinant sums of
    semistoch.f90:2044-2060 and the k loop of hci.f90:1616-1632 against the oracle; eps_pt_big chosen so that both the
    'big' and the 'small' elements occur"""
    import sqmc_b200 as sq
    s, r = c2_space_ts if ts else c2_space
    up, dn, w = label_sorted(r)
    H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP, time_sym=ts, z=1))
    for n_mc, eps_pt, eps_big, seed in [(50, 1e-5, 2e-4, 1), (200, 1e-6, 5e-5, 2), (2, 1e-5, 1e-3, 3)]:
        idx, wop = _draw_sample(w, n_mc, seed)
        e, nc = H.second_order_pt_sample(up, dn, up[idx], dn[idx], w[idx], wop, n_mc, r["energy"][0], eps_pt, eps_big)
        oe, onc = s.pt2_sample(up, dn, up[idx], dn[idx], w[idx], wop, n_mc, r["energy"][0], eps_pt, eps_big)
        assert nc == onc and nc > len(idx)
        assert abs(e - oe) <= 1e-12 * max(1.0, abs(oe) / 1e-3) and e != 0.0
    # eps_pt_big = eps_pt: every kept element is "big" unless it equals the threshold -> the sample energy (almost) vanishes
    idx, wop = _draw_sample(w, 100, 4)
    e0, _ = H.second_order_pt_sample(up, dn, up[idx], dn[idx], w[idx], wop, 100, r["energy"][0], 1e-5, 1e-5)
    oe0, _ = s.pt2_sample(up, dn, up[idx], dn[idx], w[idx], wop, 100, r["energy"][0], 1e-5, 1e-5)
    assert abs(e0 - oe0) < 1e-14


def test_pt_sample_rejects_bad_arguments(c2_space):
    import sqmc_b200 as sq
    s, r = c2_space
    up, dn, w = label_sorted(r)
    H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP))
    with pytest.raises(Exception, match="n_mc"):
        H.second_order_pt_sample(up, dn, up[:3], dn[:3], w[:3], np.ones(3), 1, r["energy"][0], 1e-5, 1e-4)
    with pytest.raises(Exception, match="non-zero"):
        H.second_order_pt_sample(up, dn, up[:3], dn[:3], np.zeros(3), np.ones(3), 10, r["energy"][0], 1e-5, 1e-4)
    with pytest.raises(Exception, match="odd"):
        H.second_order_pt_alias(up, dn, w, r["energy"][0], 1e-5, 1e-4, 10, 1e-3, [1, 2, 3, 4])
