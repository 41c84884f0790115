"""Host-side logic and the C-ABI surface (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import C2_FCIDUMP, C2_ORBSYM, ROOT, have_gpu


def test_library_exports_every_declared_symbol():
    from sqmc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sqmc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sqmc_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = C.CDLL(_lib.so_path())
    for name in declared:
        assert hasattr(L, name), "libsqmc_b200.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_gpu():
    import sqmc_b200 as sq
    with pytest.raises(sq.SqmcError) as e:
        sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49))
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "sqmc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), "%s mentions the oracle" % f


def test_chem_setup_matches_oracle_tables(oracle):
    import sqmc_b200 as sq
    cs = sq.ChemSystem(C2_FCIDUMP)
    t = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM).chem_tables()
    assert np.array_equal(cs.integrals, t["integrals"])
    assert np.array_equal(cs.combine_2.ravel(order="F"), t["combine_2"])
    assert cs.enuc == t["enuc"] and cs.hf_up == t["hf_up"] and cs.hf_dn == t["hf_dn"]
    assert np.array_equal(cs.orbital_symmetries, t["orbital_symmetries"])
    assert cs.norb == 26 and cs.nup == 4 and cs.ndn == 4
    assert len(cs.integrals) == 71631


def test_all_nine_geometries_load():
    import sqmc_b200 as sq
    for r in ("1.0", "1.1", "1.2", "1.24253", "1.3", "1.4", "1.6", "1.8", "2.0"):
        cs = sq.ChemSystem(os.path.join(ROOT, "data", "C2_v2z_curve", "r" + r, "FCIDUMP"))
        assert cs.norb == 26 and sorted(cs.orb_order[:26].tolist()) == list(range(1, 27))


def test_heg_and_hubbard_setup_match_oracle(oracle):
    import sqmc_b200 as sq
    for cut in (1.49, 2.0):
        hs = sq.HegSystem(3, 0.5, 14, 7, cut)
        ho = oracle.System.heg(3, 0.5, 14, 7, cut).heg_tables()
        assert np.array_equal(hs.k_vectors, ho["k_vectors"]) and hs.length_cell == ho["length_cell"]
    hu = sq.HubbardKSystem(4, 4, 1.0, 4.0, 8, 8)
    huo = oracle.System.hubbardk(4, 4, 1.0, 4.0, 8, 8).hubbardk_tables()
    assert np.array_equal(hu.k_vectors, huo["k_vectors"]) and np.array_equal(hu.k_energies, huo["k_energies"])
    assert hu.ubyn == huo["ubyn"] == 0.25


def test_c2_space_generator(oracle):
    import sqmc_b200 as sq
    from sqmc_b200 import spaces
    cs = sq.ChemSystem(C2_FCIDUMP)
    up, dn, total = spaces.c2_lowest_energy_space(cs, 5000)
    assert total == 27944940  # A_g sector of C2 cc-pVDZ frozen core
    assert len(up) == 5000 and up.dtype == np.uint64
    lab = list(zip(up[:, 0].tolist(), dn[:, 0].tolist()))
    assert lab == sorted(lab) and len(set(lab)) == 5000
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM)
    d = S.elements(up, dn, up, dn)
    up2, dn2, _ = spaces.c2_lowest_energy_space(cs, 6000)
    d2 = S.elements(up2, dn2, up2, dn2)
    assert d.max() <= np.sort(d2)[5000:].min() + 1e-9   # the 5000 chosen are the lowest
    hf = (int(up[0, 0]), int(dn[0, 0]))
    assert hf == (cs.hf_up, cs.hf_dn)
    upt, dnt, tot_t = spaces.c2_lowest_energy_space(cs, 1000, time_sym=True)
    assert tot_t == 13979945 and np.all(upt[:, 0] <= dnt[:, 0])


def test_partition_rows_rule():
    from sqmc_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(5)
    for n, R in ((1000, 2), (1000, 8), (7, 8), (1, 2)):
        w = rng.integers(1, 2000, n).astype(np.int64)
        prefix = np.zeros(n + 1, dtype=np.int64)
        prefix[1:] = np.cumsum(w)
        starts = np.zeros(R + 1, dtype=np.int64)
        assert L.sqmc_b200_partition_rows(prefix.ctypes.data_as(C.c_void_p), n, R, starts.ctypes.data_as(C.c_void_p)) == 0
        assert starts[0] == 0 and starts[-1] == n and np.all(np.diff(starts) >= 0)
        if n >= 100 * R:
            loads = np.array([prefix[starts[r + 1]] - prefix[starts[r]] for r in range(R)])
            assert loads.max() - loads.min() <= 2 * w.max()


def test_splitmix_vector_is_rank_independent():
    from sqmc_b200 import spaces
    x = spaces.splitmix_vector(1000)
    assert abs(np.linalg.norm(x) - 1) < 1e-14 and np.array_equal(x, spaces.splitmix_vector(1000))
    assert x.min() < 0 < x.max()


def test_reference_input_files_are_parsed():
    """sqmc_b200.hci.read_input on the two input layouts the reference ships"""
    from sqmc_b200 import hci
    c2 = hci.read_input(os.path.join(ROOT, "data", "C2_v2z_curve", "r1.24253", "i_1sigma_g"))
    assert c2["hamiltonian_type"] == "chem" and (c2["nelec"], c2["nup"], c2["norb"]) == (8, 4, 26)
    assert c2["time_sym"] is True and c2["z"] == 1 and c2["n_states"] == 2
    assert (c2["eps_var"], c2["eps_pt"]) == (1e-3, 1e-7) and c2["eps_var_sched"] == [2e-3, 2e-3]     # 2*2e-3 repeat syntax
    assert c2["orbital_symmetries"] == [1, 5, 3, 2, 1, 7, 6, 5, 1, 2, 3, 1, 6, 7, 5, 4, 1, 5, 3, 2, 8, 5, 1, 7, 6, 5]
    assert c2["hf_symmetry"] == 1 and c2["n_mc"] == 0
    heg = hci.read_input(os.path.join(ROOT, "tests", "golden", "heg_i_det"))
    assert heg["hamiltonian_type"] == "heg" and (heg["n_dim"], heg["r_s"], heg["nelec"], heg["nup"]) == (3, 0.5, 14, 7)
    assert heg["cutoff_radius"] == 1.49 and (heg["eps_var"], heg["eps_pt"], heg["n_states"]) == (1e-3, 2e-7, 1)
    s = hci.eps_schedule(1e-3, [2e-3, 2e-3])
    assert len(s) == 30 and s[0] == s[1] == 2e-3 and s[2] == s[29] == 1e-3


class _OracleBackedHamiltonian:
    """Duck-typed stand-in for SparseHamiltonian whose heavy steps are the CPU oracle: lets the host-side driver loop
    (sqmc_b200.hci.perform_hci) be checked without a GPU.  Test infrastructure only."""

    def __init__(self, S, O):
        self.S, self.O = S, O
        self.mat = None

    def diagonal(self, up, dn):
        return self.S.elements(up, dn, up, dn)

    def get_next_det_list(self, up, dn, coeffs, min_h, eps):
        return self.S.select(up, dn, coeffs, min_h, eps)

    def generate_sparse_ham_upper_triangular(self, up, dn, ndet_old=0):
        self.mat = self.S.build_upper(up, dn)
        return len(self.mat[1])

    def davidson_sparse(self, n_states=1, initial_vector=None):
        return self.O.davidson(*self.mat, n_states=n_states, v0=initial_vector)

    def second_order_pt(self, up, dn, wts, e, eps_pt):
        return self.S.pt2(up, dn, wts, e, eps_pt)


def test_perform_hci_driver_loop_reproduces_reference_log(oracle):
    """the Python mirror of perform_hci (schedule, selection coefficients, stopping rules, PT call) driven by the oracle:
    the reference's HEG end-to-end input must give its golden log (o_det_ref:261-437)"""
    import json
    import types
    from sqmc_b200 import hci
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "heg_o_det_ref.json")))
    cfg = hci.read_input(os.path.join(ROOT, "tests", "golden", "heg_i_det"))
    S = oracle.System.heg(cfg["n_dim"], cfg["r_s"], cfg["nelec"], cfg["nup"], cfg["cutoff_radius"])
    system = types.SimpleNamespace(hf_up=(1 << cfg["nup"]) - 1, hf_dn=(1 << (cfg["nelec"] - cfg["nup"])) - 1, time_sym=False)
    res = hci.perform_hci(_OracleBackedHamiltonian(S, oracle), system, cfg["eps_var"], cfg["eps_var_sched"], n_states=cfg["n_states"],
                          eps_pt=cfg["eps_pt"])
    assert [it["n_det"] for it in res["iterations"]][:2] == gold["n_det"] and [it["nnz"] for it in res["iterations"]][:2] == gold["nnz"]
    assert len(res["up"]) == gold["pt"]["ndets"] and abs(res["energy"][0] - gold["pt"]["variational_energy"]) < 5e-9
    de, nconn = res["pt"][0]
    assert nconn == gold["pt"]["ndets_connected"] and abs(res["energy"][0] + de - gold["pt"]["total_energy"]) < 1e-9


def test_perform_hci_driver_loop_two_states_matches_oracle_loop(oracle):
    """n_states = 2 with an eps_var schedule (the shipped C2 input): the Python loop and the oracle's own perform_hci
    restatement must walk through the same determinant counts and energies"""
    import json
    import types
    from conftest import C2_FCIDUMP, C2_ORBSYM
    from sqmc_b200 import hci
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "c2_s1_hci.json")))["runs"]["n_states=2"]
    S = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=True, z=1, hf_symmetry=1)
    tab = S.chem_tables()   # the HF determinant in the reordered orbital numbering (chemistry.f90:694-805)
    system = types.SimpleNamespace(hf_up=tab["hf_up"], hf_dn=tab["hf_dn"], time_sym=True)
    res = hci.perform_hci(_OracleBackedHamiltonian(S, oracle), system, 1e-3, [2e-3, 2e-3], n_states=2)
    assert [it["n_det"] for it in res["iterations"]] == gold["n_det"]
    assert np.max(np.abs(np.array([it["energy"] for it in res["iterations"]]) - np.array(gold["iter_energy"]))) < 1e-8


def test_perform_hci_cycles_when_selection_finds_nothing_new(oracle):
    """hci.f90:413-417: an iteration whose selection adds no determinant (a repeated eps_var in the schedule) is cycled --
    no H build, no Davidson, no extra iteration entry -- before the exit tests are looked at"""
    import types
    from sqmc_b200 import hci
    S = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    system = types.SimpleNamespace(hf_up=(1 << 7) - 1, hf_dn=(1 << 7) - 1, time_sym=False)

    class Counting(_OracleBackedHamiltonian):
        builds = 0
        empty_selections = 0

        def get_next_det_list(self, up, dn, coeffs, min_h, eps):
            r = super().get_next_det_list(up, dn, coeffs, min_h, eps)
            Counting.empty_selections += int(len(r[0]) == 0)
            return r

        def generate_sparse_ham_upper_triangular(self, up, dn, ndet_old=0):
            Counting.builds += 1
            return super().generate_sparse_ham_upper_triangular(up, dn, ndet_old=ndet_old)

    log = []
    res = hci.perform_hci(Counting(S, oracle), system, 1e-3, [4e-3] * 6, n_states=1, log=log.append)
    assert Counting.empty_selections >= 1, "the schedule was meant to repeat an eps_var until nothing new is found"
    assert any("Cycling hci iteration" in ln for ln in log)
    assert Counting.builds == len(res["iterations"])          # cycled iterations built nothing and logged nothing
    counts = [it["n_det"] for it in res["iterations"]]
    assert counts == sorted(set(counts))                       # every logged iteration grew the list
    # same final space and energy as the oracle's own perform_hci restatement with the same schedule
    ref = S.hci(1e-3, eps_var_sched=[4e-3] * 6, n_states=1)
    assert len(res["up"]) == len(ref["up"]) and abs(res["energy"][0] - ref["energy"][0]) < 1e-9


def test_perform_hci_checkpoint_dump_and_resume(oracle, tmp_path):
    """dump_wf_var: the variational stage writes wf_eps_var=1.00E-3 (label-sorted, hci.f90:569-625); a second run finds it,
    skips the stage (:194-231) and obtains the same PT correction"""
    import json
    import types
    from sqmc_b200 import formats, hci
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "heg_o_det_ref.json")))
    S = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    system = types.SimpleNamespace(hf_up=(1 << 7) - 1, hf_dn=(1 << 7) - 1, time_sym=False)
    H = _OracleBackedHamiltonian(S, oracle)
    first = hci.perform_hci(H, system, 1e-3, n_states=1, eps_pt=8.192e-4, wf_dir=str(tmp_path), dump_wf_var=True)
    path = tmp_path / "wf_eps_var=1.00E-3"
    assert path.exists() and not first["from_checkpoint"]
    ck = formats.read_wf(path)
    lab = [(int(u[1]) << 64 | int(u[0]), int(d[1]) << 64 | int(d[0])) for u, d in zip(ck["up"], ck["dn"])]
    assert lab == sorted(lab) and len(lab) == 9475
    calls = {"select": 0}
    orig = H.get_next_det_list
    H.get_next_det_list = lambda *a: (calls.__setitem__("select", calls["select"] + 1), orig(*a))[1]
    second = hci.perform_hci(H, system, 1e-3, n_states=1, eps_pt=8.192e-4, wf_dir=str(tmp_path), dump_wf_var=True)
    assert second["from_checkpoint"] and calls["select"] == 0 and second["iterations"] == []
    assert second["pt"][0] == first["pt"][0]
    assert second["pt"][0][1] == gold["pt_big"]["ndets_connected"] and abs(second["pt"][0][0] - gold["pt_big"]["pt_correction"]) < 5e-10


def test_driver_output_passes_the_reference_e2e_checker(oracle):
    """src/e2e_tests/e2e_check.py greps 'Variational energy=' and 'Second-order PT energy lowering=' from a run's output and
    compares them with o_det_ref (1 % tolerance): the same regular expressions applied to the driver's log"""
    import json
    import re
    import types
    from sqmc_b200 import hci
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "heg_o_det_ref.json")))["pt"]
    S = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    system = types.SimpleNamespace(hf_up=(1 << 7) - 1, hf_dn=(1 << 7) - 1, time_sym=False)
    lines = []
    hci.perform_hci(_OracleBackedHamiltonian(S, oracle), system, 1e-3, n_states=1, eps_pt=2e-7, log=lines.append)
    text = "\n".join(lines)
    ev = float(re.search(r"Variational energy.*=\s*([-+]?[0-9]*\.?[0-9]+)", text).group(1))
    pt = float(re.search(r"Second-order PT energy lowering.*=\s*([-+]?[0-9]*\.?[0-9]+)", text).group(1))
    assert "Variational energy=                   58.276906085" in text          # o_det_ref:434, character for character
    assert "Second-order PT energy lowering=      -0.000939196" in text          # o_det_ref:435
    assert abs(ev - gold["variational_energy"]) < 1e-9 and abs(pt - gold["pt_correction"]) < 1e-9
