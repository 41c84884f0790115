"""The COMPILED host side (host/sqmc_b200_host.hpp + host/dropin_demo.cpp): a C++ caller that uses the C ABI the way the
Fortran driver would (same routine names / argument meaning as the reference), checked against the oracle."""
import itertools
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import C2_FCIDUMP, ROOT

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "host")


def _build_demo():
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return os.path.join(HOST, "dropin_demo")


def _read_out(path):
    out = []
    with open(path, "rb") as f:
        for dt in (np.int64, np.int64, np.float64, np.float64, np.float64, np.float64, np.float64, np.float64):
            (n,) = struct.unpack("q", f.read(8))
            out.append(np.frombuffer(f.read(8 * n), dtype=dt).copy())
    return out


def _run(tmp_path, header, blobs):
    exe = _build_demo()
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.asarray(header, dtype=np.int64).tobytes())
        for b in blobs:
            f.write(np.ascontiguousarray(b).tobytes())
    p = subprocess.run([exe, inp, outp], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return _read_out(outp)


def test_cpp_caller_chem(oracle, c2_space_ts, tmp_path):
    import sqmc_b200 as sq
    s, r = c2_space_ts
    cs = sq.ChemSystem(C2_FCIDUMP, time_sym=True, z=1)
    n = len(r["up"])
    x = np.random.default_rng(5).uniform(-1, 1, n)
    hdr = [0, cs.norb, cs.nup, cs.ndn, 1, 1, 0, 0, len(cs.integrals), n, 1, 0]
    cnt, idx, val, y, evals, evecs, ritz, dw = _run(tmp_path, hdr, [cs.integrals, np.asfortranarray(cs.combine_2).ravel(order="F").astype(np.int32),
                                                                   r["up"], r["dn"], x])
    ref = s.build_upper(r["up"], r["dn"])
    assert np.array_equal(cnt, ref[0]) and np.array_equal(idx, ref[1]) and np.array_equal(val, ref[2])
    yref = oracle.matvec_upper(*ref, x)
    assert np.max(np.abs(y - yref)) <= 1e-12 * np.max(np.abs(yref))
    dref = oracle.davidson(*ref, n_states=1)
    assert abs(evals[0] - dref["evals"][0]) < 1e-8 and np.max(np.abs(ritz - dref["ritz"].ravel())) < 1e-8
    _, dwr = oracle.projector_step(ref[0], ref[1], -0.01 * ref[2], 0.01, float(evals[0]), x)
    assert np.max(np.abs(dw - dwr)) <= 1e-12 * np.max(np.abs(dwr))


def test_cpp_caller_hubbard_hf_to_psit(oracle, tmp_path):
    import sqmc_b200 as sq
    hs = sq.HubbardKSystem(4, 4, 1.0, 4.0, 3, 3)
    so = oracle.System.hubbardk(4, 4, 1.0, 4.0, 3, 3)
    strings = [sum(1 << o for o in c) for c in itertools.combinations(range(16), 3)]
    dets = sorted((u, d) for u in strings for d in strings if hs.total_momentum(u, d) == (0, 0))
    up = oracle.dets_to_u64([u for u, d in dets])
    dn = oracle.dets_to_u64([d for u, d in dets])
    n = len(dets)
    x = np.random.default_rng(6).uniform(-1, 1, n)
    hdr = [2, 16, 3, 3, 0, 1, 4, 4, 0, n, 1, 1]
    cnt, idx, val, y, evals, evecs, ritz, dw = _run(tmp_path, hdr, [hs.k_vectors, hs.k_energies, np.array([hs.ubyn]), up, dn, x])
    ref = so.build_upper(up, dn, hf_to_psit=True)
    assert ref[0][0] == 1 and ref[2][0] == 0.0           # first row: a single zero diagonal entry (hubbard.f90:9636-9643)
    assert np.array_equal(cnt, ref[0]) and np.array_equal(idx, ref[1]) and np.array_equal(val, ref[2])
    yref = oracle.matvec_upper(*ref, x)
    assert np.max(np.abs(y - yref)) <= 1e-12 * np.max(np.abs(yref))
    # and through the Python mirror
    H = sq.SparseHamiltonian(hs)
    assert H.generate_sparse_ham_upper_triangular(up, dn, hf_to_psit=True) == len(ref[1])
    got = H.export_upper()
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def test_cpp_caller_semistochastic_pt(oracle, tmp_path):
    """host/pt_demo.cpp = the semistochastic branch of do_pt (hci.f90:4245-4300) as a compiled caller: second_order_pt with
    eps_pt_big, then second_order_pt_alias with the caller's rannyu state, on the reference's HEG end-to-end case
    (src/e2e_tests/heg/o_st_ref:432,873): -0.000199339, -0.000729402 +- 0.000009966 after 143 samples, total lowering -0.000928741"""
    import json
    import sqmc_b200 as sq
    from conftest import label_sorted
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "heg_o_det_ref.json")))
    st, big = g["pt_stochastic"], g["pt_big"]
    S = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    r = S.hci(1e-3, n_states=1)
    up, dn, w = label_sorted(r)
    hs = sq.HegSystem(3, 0.5, 14, 7, 1.49)
    subprocess.check_call(["make", "-C", HOST, "-s"])
    seed = list(st["irand_seed_1"]); seed[3] = 2 * (seed[3] // 2) + 1          # setrn (rannyu.f90:19)
    hdr = [hs.norb, hs.n_dim, hs.nup, hs.ndn, len(up), st["n_mc"], 1000] + seed + [0]
    par = np.array([hs.length_cell, r["energy"][0], st["eps_pt"], st["eps_pt_big"], st["target_error"]])
    inp, outp = str(tmp_path / "pt_in.bin"), str(tmp_path / "pt_out.bin")
    with open(inp, "wb") as f:
        f.write(np.asarray(hdr, dtype=np.int64).tobytes())
        for b in (par, hs.k_vectors, up, dn, w):
            f.write(np.ascontiguousarray(b).tobytes())
    p = subprocess.run([os.path.join(HOST, "pt_demo"), inp, outp], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    out = np.fromfile(outp, dtype=np.float64)
    pt_big, nconn_big, pt_diff, sd, n_samples = out[0], int(out[1]), out[2], out[3], int(out[4])
    assert nconn_big == big["ndets_connected"] and abs(pt_big - big["pt_correction"]) < 5.1e-10
    assert n_samples == len(st["samples"]) == 143
    assert abs(pt_diff - st["pt_diff"]) < 5.1e-10 and abs(sd - st["std_dev"]) < 5.1e-10
    assert abs(pt_big + pt_diff - st["pt_total"]) < 1.1e-9
    assert int(out[6]) % 2 == 1                                                # the advanced rannyu state stays odd
    assert "Second-order PT energy lowering=" in p.stdout
