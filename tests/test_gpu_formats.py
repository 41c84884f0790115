"""File-format interop through the GPU path: a deterministic projector dumped in the reference's text format and loaded
back serves the same walk step; a wf_eps_var checkpoint resumes an HCI run."""
import numpy as np
import pytest

from conftest import C2_FCIDUMP

pytestmark = pytest.mark.gpu


def test_dtm_projector_file_round_trip_serves_the_same_step(oracle, heg_space, tmp_path):
    import sqmc_b200 as sq
    s, r = heg_space
    up, dn = r["up"], r["dn"]
    n = len(up)
    hs = sq.HegSystem(3, 0.5, 14, 7, 1.49)
    H = sq.SparseHamiltonian(hs)
    H.generate_sparse_ham_upper_triangular(up, dn)
    tau, e_trial = 0.02, float(r["energy"][0])
    w = np.abs(r["wts"][:, 0]) + 1e-3
    path = tmp_path / "dtm_projector.out"
    H.dump_dtm_projector(path, up, dn, dtm_energy=e_trial)          # the file holds H itself
    H.scale_values(-tau)
    dw = H.projector_step(tau, e_trial, w)
    path2 = tmp_path / "dtm_projector_scaled.out"
    H.dump_dtm_projector(path2, up, dn, dtm_energy=e_trial, tau=tau)  # from the -tau*H matrix: the same file up to rounding of /tau
    G = sq.SparseHamiltonian(hs)
    fu, fd = G.load_dtm_projector(path, tau, nup=7, ndn=7)            # reads, imports, multiplies by -tau
    assert np.array_equal(fu, up) and np.array_equal(fd, dn)
    assert G.nnz()["nnz_upper"] == H.nnz()["nnz_upper"] == 165193     # golden size, o_det_ref:330
    dw2 = G.projector_step(tau, e_trial, w)
    assert np.max(np.abs(dw2 - dw)) <= 1e-13 * np.max(np.abs(dw))
    cnt, idx, val = s.build_upper(up, dn)
    _, dw_ref = oracle.projector_step(cnt, idx, -tau * val, tau, e_trial, w)
    assert np.max(np.abs(dw2 - dw_ref)) <= 1e-12 * np.max(np.abs(dw_ref))
    from sqmc_b200 import formats
    a, b = formats.read_dtm_projector(path, 7, 7), formats.read_dtm_projector(path2, 7, 7)
    assert np.array_equal(a["values"], val)                            # bit-exact H through the text file
    assert np.max(np.abs(a["values"] - b["values"])) <= 4e-16 * np.max(np.abs(val))


def test_wf_checkpoint_resumes_an_hci_run(tmp_path):
    import sqmc_b200 as sq
    from sqmc_b200 import formats, spaces
    cs = sq.ChemSystem(C2_FCIDUMP)
    H = sq.SparseHamiltonian(cs)
    log = []
    up, dn, wts, e = spaces.hci_space(H, cs, 10**9, eps_schedule=(1e-3, 3e-4), log=log)
    path = tmp_path / formats.wf_filename(3e-4)
    assert path.name == "wf_eps_var=3.00E-4"
    formats.write_wf(path, up, dn, wts, [e])
    ck = formats.read_wf(path)
    assert np.array_equal(ck["up"], up) and np.array_equal(ck["dn"], dn) and ck["energies"][0] == e
    # a fresh process would read the checkpoint, rebuild H for its determinants and restart Davidson from its weights
    G = sq.SparseHamiltonian(cs)
    G.generate_sparse_ham_upper_triangular(ck["up"], ck["dn"])
    d = G.davidson_sparse(n_states=1, initial_vector=ck["wts"])
    assert abs(d["evals"][0] - e) < 1e-9 and d["n_matvec"] <= 4
    # and the next selection step continues exactly where the first run would have
    mh = np.full(len(up), 9e99)
    a = H.get_next_det_list(up, dn, np.abs(wts[:, 0]), mh.copy(), 1e-4)
    b = G.get_next_det_list(ck["up"], ck["dn"], np.abs(ck["wts"][:, 0]), mh.copy(), 1e-4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and len(a[0]) == 171060 - 42456
