"""The reference's own input files run end to end through the GPU path (sqmc_b200.hci.run_input): variational loop + PT."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_heg_e2e_input_reproduces_the_reference_output():
    """src/e2e_tests/heg/i_det -> o_det_ref: 277 / 9475 determinants, E_var 58.276906085, 501881 connected determinants,
    PT -0.000939196, total 58.275966889"""
    from sqmc_b200 import hci
    gold = json.load(open(os.path.join(HERE, "golden", "heg_o_det_ref.json")))
    res = hci.run_input(os.path.join(HERE, "golden", "heg_i_det"), log=None)
    assert [it["n_det"] for it in res["iterations"]][:2] == gold["n_det"]
    assert [it["nnz"] for it in res["iterations"]][:2] == gold["nnz"]
    assert len(res["up"]) == gold["pt"]["ndets"]
    assert abs(res["energy"][0] - gold["pt"]["variational_energy"]) < 5e-9
    de, nconn = res["pt"][0]
    assert nconn == gold["pt"]["ndets_connected"] and abs(de - gold["pt"]["pt_correction"]) < 5e-10
    assert abs(res["energy"][0] + de - gold["pt"]["total_energy"]) < 1e-9


def test_c2_shipped_input_matches_the_oracle_run(oracle, c2_hci_full):
    """C2_v2z_curve/r1.24253/i_1sigma_g as shipped (time_sym = t, n_states = 2, eps_var_sched = 2*2e-3, eps_pt = 1e-7):
    variational stage against the committed oracle fixture, PT of both states against the oracle"""
    from sqmc_b200 import hci
    gold = json.load(open(os.path.join(HERE, "golden", "c2_s1_hci.json")))["runs"]["n_states=2"]
    res = hci.run_input(os.path.join(ROOT, "data", "C2_v2z_curve", "r1.24253", "i_1sigma_g"), log=None)
    assert [it["n_det"] for it in res["iterations"]] == gold["n_det"]
    assert [it["nnz"] for it in res["iterations"]] == gold["nnz"]
    assert np.max(np.abs(np.array([it["energy"] for it in res["iterations"]]) - np.array(gold["iter_energy"]))) < 1e-8
    assert np.max(np.abs(res["energy"] - np.array(gold["energy"]))) < 1e-8
    S, _ = c2_hci_full
    for s in range(2):
        ode, onc = S.pt2(res["up"], res["dn"], res["wts"][:, s], res["energy"][s], 1e-7)
        de, nconn = res["pt"][s]
        assert nconn == onc and abs(de - ode) < 1e-10 and de < 0
