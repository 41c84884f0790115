#!/usr/bin/env python
"""Regenerates the committed golden fixtures.

  heg_o_det_ref.json   transcribed by regex from the REFERENCE's own golden log
                       /root/reference/src/e2e_tests/heg/o_det_ref (only readable in the build
                       container; the GPU box never needs it).
  c2_s1_hci.json       oracle HCI on the shipped C2 cc-pVDZ input (C2_v2z_curve/r1.24253/i_1sigma_g).
                       The reference ships no output for this input: "parity unpinned by the
                       reference", pinned by the oracle (itself pinned on the HEG log).
  c2_hci_sched.json    oracle HCI with the benchmark's eps_var schedule (1e-3, 3e-4, 1e-4; one iteration each) on
                       C2 r1.24253 time_sym=f: sizes, nnz, energies and a SHA-256 of the final determinant list.
  c2_small_space.npz   2000 lowest-energy A_g determinants of C2 + the oracle's upper-triangular H.

  heg_i_det            INPUT DATA, not code: the reference's end-to-end test input src/e2e_tests/heg/i_det, kept verbatim
                       (like the FCIDUMPs under data/) so that the run which must reproduce o_det_ref reads the very
                       same file; refreshed by this script when the reference is present.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def heg_from_reference_log():
    path = "/root/reference/src/e2e_tests/heg/o_det_ref"
    if not os.path.exists(path):
        print("reference log not present; keeping the committed heg_o_det_ref.json")
        return
    lines = open(path).read().splitlines()
    out = {"source": "src/e2e_tests/heg/o_det_ref", "input": {"n_dim": 3, "r_s": 0.5, "nelec": 14, "nup": 7, "cutoff_radius": 1.49, "eps_var": 1e-3}}
    kp = []
    for i, ln in enumerate(lines):
        if ln.strip() == "K Points":
            j = i + 1
            while re.match(r"\s*\d+\s+-?\d", lines[j]):
                t = lines[j].split()
                kp.append([float(x) for x in t[1:5]])
                j += 1
            out["k_points_line"] = i + 1
            break
    out["k_points"] = kp
    for i, ln in enumerate(lines):
        m = re.search(r"HF kinetic, exchange, total energies =\s+(\S+)\s+(\S+)\s+(\S+)", ln)
        if m:
            out["hf_energy"] = float(m.group(3)); out["hf_energy_line"] = i + 1
    out["n_det"], out["nnz"], out["ritz"], out["davidson_final"] = [], [], [], []
    cur = None
    for i, ln in enumerate(lines):
        m = re.search(r"n_det=\s*(\d+).*# of nonzero elem in H=\s*(\d+)", ln)
        if m:
            out["n_det"].append(int(m.group(1))); out["nnz"].append(int(m.group(2)))
            cur = []; out["ritz"].append(cur)
        m = re.search(r"Iteration, Eigenvalues=\s*\d+\s+(\S+)", ln)
        if m and cur is not None:
            cur.append(float(m.group(1)))
        m = re.search(r"davidson_sparse: n, iter, Lowest eigenvalue =\s*(\d+)\s+(\d+)\s+(\S+)", ln)
        if m:
            out["davidson_final"].append({"n": int(m.group(1)), "iter": int(m.group(2)), "energy": float(m.group(3))})
    for i, ln in enumerate(lines):
        m = re.search(r"PT_correction, eps_pt, ndets_connected for fully deterministic run=\s*(\S+)\s+(\S+)\s+(\d+)", ln)
        if m:
            out["pt"] = {"pt_correction": float(m.group(1)), "eps_pt": float(m.group(2)), "ndets_connected": int(m.group(3)), "line": i + 1}
        m = re.search(r"ndets, ndets_connected\(total\), Variational, PT, Total Energies=\s*(\d+)\s+(\d+)\s+(\S+)\s+(\S+)\s+(\S+)", ln)
        if m:
            out["pt"].update({"ndets": int(m.group(1)), "variational_energy": float(m.group(3)), "total_energy": float(m.group(5)), "total_line": i + 1})
    coefs = []
    for i, ln in enumerate(lines):
        if ln.startswith("Final variational wavefunctions"):
            for j in range(i + 1, i + 21):
                t = lines[j].split()
                coefs.append({"up": int(t[1]), "dn": int(t[2]), "coef": float(t[3])})
            break
    out["final_coefficients"] = coefs
    # second golden log of the same system: src/e2e_tests/heg/o_st_ref (semistochastic PT; its deterministic first stage)
    st = "/root/reference/src/e2e_tests/heg/o_st_ref"
    if os.path.exists(st):
        for i, ln in enumerate(open(st).read().splitlines()):
            m = re.search(r"PT_correction, eps_pt_big, ndets_connected for short deterministic run=\s*(\S+)\s+(\S+)\s+(\d+)", ln)
            if m:
                out["pt_big"] = {"source": "src/e2e_tests/heg/o_st_ref", "pt_correction": float(m.group(1)), "eps_pt_big": float(m.group(2)),
                                 "ndets_connected": int(m.group(3)), "line": i + 1}
        # the stochastic stage of the same log (second_order_pt_alias, hci.f90:1314): seeds, n_mc, every printed sample
        stl = open(st).read().splitlines()
        sto = {"source": "src/e2e_tests/heg/o_st_ref", "samples": [], "n_ref": []}
        in_alias = False
        for i, ln in enumerate(stl):
            m = re.match(r"random number seeds \[mas(\d{4})(\d{4})(\d{4})(\d{4}) ", ln)
            if m:
                sto["irand_seed_1"] = [int(m.group(k)) for k in range(1, 5)]          # do_walk.f90:231-238 (4i4,x,4i4), first set
            m = re.search(r"eps_var, eps_pt, eps_pt_big, target_error=\s*(\S+)\s+(\S+)\s+(\S+)\s+(\S+)", ln)
            if m:
                sto["target_error"] = float(m.group(4))
            m = re.search(r"single-list stochastic method with N_MC=\s*(\d+)", ln)
            if m:
                sto["n_mc"] = int(m.group(1)); in_alias = True
            m = re.match(r"n_connected_dets,n_ref=\s*\d+\s+(\d+)", ln)
            if m and in_alias:
                sto["n_ref"].append(int(m.group(1)))
            m = re.match(r"Sample, E_2pt_now, E_2pt estimate, total energy=\s*(\d+)\s+(\S+)\s+(\S+)\s+(\S+) \+-\s+(\S+)", ln)
            if m:
                sto["samples"].append({"sample": int(m.group(1)), "e_now": float(m.group(2)), "estimate": float(m.group(3)),
                                       "total": float(m.group(4)), "line": i + 1})
            m = re.search(r"Second-order PT energy lowering=\s*(\S+) \+- (\S+) \(\s*(\S+)\s+(\S+)\)", ln)
            if m:
                sto["pt_total"], sto["std_dev"], sto["pt_diff"] = float(m.group(1)), float(m.group(2)), float(m.group(4))
            m = re.search(r"ndets, ndets_connected\(total\), Variational, PT, Total Energies=\s*(\d+)\s+(\d+)", ln)
            if m:
                sto["ndets_connected_last_sample"] = int(m.group(2))                  # ndets_connected left by the last find_doubly_excited call
        sto["eps_pt"], sto["eps_pt_big"] = 2e-7, out["pt_big"]["eps_pt_big"]          # i_st: pt_eps; &selected_ci eps_pt_big
        out["pt_stochastic"] = sto
    json.dump(out, open(os.path.join(HERE, "heg_o_det_ref.json"), "w"), indent=1)
    print("wrote heg_o_det_ref.json:", out["n_det"], out["nnz"])


def c2_from_oracle():
    from conftest import C2_FCIDUMP, C2_ORBSYM
    from oracle import oracle as O
    import sqmc_b200 as sq
    from sqmc_b200 import spaces
    res = {"source": "oracle (parity unpinned by the reference: no output shipped for C2_v2z_curve)",
           "input": "data/C2_v2z_curve/r1.24253/i_1sigma_g: eps_var=1e-3, eps_var_sched=2*2e-3, time_sym=t, z=1, hf_symmetry=1", "runs": {}}
    for n_states in (1, 2):
        S = O.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=True, z=1, hf_symmetry=1)
        r = S.hci(1e-3, eps_var_sched=[2e-3, 2e-3], n_states=n_states)
        res["runs"]["n_states=%d" % n_states] = {"n_det": r["ndet"].tolist(), "nnz": r["nnz"].tolist(),
                                                 "iter_energy": r["iter_energy"].tolist(), "energy": r["energy"].tolist()}
    S = O.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=False, z=1, hf_symmetry=1)
    r = S.hci(1e-3, eps_var_sched=[2e-3, 2e-3], n_states=1)
    res["runs"]["time_sym=f n_states=1"] = {"n_det": r["ndet"].tolist(), "nnz": r["nnz"].tolist(),
                                            "iter_energy": r["iter_energy"].tolist(), "energy": r["energy"].tolist()}
    json.dump(res, open(os.path.join(HERE, "c2_s1_hci.json"), "w"), indent=1)
    chem = sq.ChemSystem(C2_FCIDUMP)
    up, dn, _ = spaces.c2_lowest_energy_space(chem, 2000)
    cnt, idx, val = S.build_upper(up, dn)
    np.savez_compressed(os.path.join(HERE, "c2_small_space.npz"), up=up, dn=dn, counts=cnt, indices=idx, values=val)
    print("wrote c2_s1_hci.json, c2_small_space.npz", len(cnt), len(idx))


def c2_sched_from_oracle():
    import hashlib
    from conftest import C2_FCIDUMP, C2_ORBSYM
    from oracle import oracle as O
    S = O.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM)
    sched = [1e-3, 3e-4, 1e-4]
    r = S.hci(sched[-1], eps_var_sched=sched, max_iters=len(sched))
    order = np.lexsort((r["dn"][:, 0], r["up"][:, 0]))
    dig = hashlib.sha256(np.ascontiguousarray(np.stack([r["up"][order, 0], r["dn"][order, 0]], axis=1)).tobytes()).hexdigest()
    out = {"source": "oracle perform_hci restatement (hci.f90:359-517), one iteration per eps_var_sched entry",
           "input": "data/C2_v2z_curve/r1.24253/FCIDUMP, time_sym=f", "eps_var_sched": sched,
           "n_det": r["ndet"].tolist(), "nnz": r["nnz"].tolist(), "iter_energy": r["iter_energy"][:, 0].tolist(),
           "sha256_sorted_up_dn_u64": dig}
    json.dump(out, open(os.path.join(HERE, "c2_hci_sched.json"), "w"), indent=1)
    print("wrote c2_hci_sched.json", out["n_det"], out["nnz"])


def heg_input_file():
    src = "/root/reference/src/e2e_tests/heg/i_det"
    if os.path.exists(src):
        import shutil
        shutil.copyfile(src, os.path.join(HERE, "heg_i_det"))
        print("copied heg_i_det")


if __name__ == "__main__":
    heg_input_file()
    heg_from_reference_log()
    c2_from_oracle()
    c2_sched_from_oracle()
