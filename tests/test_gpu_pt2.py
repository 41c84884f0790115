"""Deterministic second-order PT on the device (csrc/select.cu: pt2) against the reference's golden log and the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import C2_FCIDUMP

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
