#!/usr/bin/env python
"""Row-sharded path on N GPUs (launched with torchrun): build, export, H.v, Davidson and projector
against the CPU oracle.  Used by tests/test_gpu_multi.py and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_multi_gpu_parity.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import sqmc_b200 as sq
    from sqmc_b200 import _lib
    from conftest import C2_FCIDUMP, C2_ORBSYM
    from oracle import oracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    obj = [_lib.get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    _lib.init(device=local, rank=rank, nranks=world, unique_id=obj[0])
    for time_sym in (False, True):
        S = O.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=time_sym, z=1, hf_symmetry=1)
        r = S.hci(1e-3, eps_var_sched=[2e-3, 2e-3], n_states=1, max_iters=3)
        cnt, idx, val = S.build_upper(r["up"], r["dn"])
        n = len(cnt)
        H = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP, time_sym=time_sym, z=1), device=local)
        nnz = H.generate_sparse_ham_upper_triangular(r["up"], r["dn"])
        assert nnz == len(idx), (nnz, len(idx))
        nloc, nnz_loc = H.local_rows()
        tot = torch.tensor([nloc, nnz_loc], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        assert tot[0].item() == n and tot[1].item() == 2 * nnz - n
        # export: this rank's rows (ascending caller index) against the oracle's rows
        perm = H.perm()
        gc, gi, gv = H.export_upper()
        # rows owned: internal rows [row0,row1) -> need row0: all-gather nloc
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([nloc], dtype=torch.int64, device="cuda"))
        row0 = int(sum(s.item() for s in sizes[:rank]))
        mine = np.sort(perm[row0:row0 + nloc])
        ptr = np.zeros(n + 1, dtype=np.int64)
        ptr[1:] = np.cumsum(cnt)
        assert np.array_equal(gc, cnt[mine])
        ri = np.concatenate([idx[ptr[i]:ptr[i + 1]] for i in mine]) if len(mine) else np.zeros(0, dtype=np.int64)
        rv = np.concatenate([val[ptr[i]:ptr[i + 1]] for i in mine]) if len(mine) else np.zeros(0)
        assert np.array_equal(gi, ri) and np.array_equal(gv, rv)
        x = np.random.default_rng(12345).uniform(-1, 1, n)
        y = H.matvec(x)
        yref = O.matvec_upper(cnt, idx, val, x)
        assert np.max(np.abs(y - yref)) <= 1e-12 * np.max(np.abs(yref))
        ref = O.davidson(cnt, idx, val, n_states=1)
        got = H.davidson_sparse(n_states=1)
        assert got["ritz"].shape == ref["ritz"].shape and np.max(np.abs(got["ritz"] - ref["ritz"])) < 1e-8
        assert abs(abs(np.dot(got["evecs"][:, 0], ref["evecs"][:, 0])) - 1) < 1e-6
        # two states: each block's pair of H.v goes through the two-vector kernel after an all-gather of the interleaved slices
        ref2 = O.davidson(cnt, idx, val, n_states=2)
        got2 = H.davidson_sparse(n_states=2)
        assert got2["n_matvec"] == ref2["n_matvec"] and np.max(np.abs(got2["evals"] - ref2["evals"])) < 1e-8
        # the whole variational loop sharded: selection replicated, build + Davidson across the ranks
        if not time_sym:
            from sqmc_b200 import spaces
            import hashlib, json
            gold = json.load(open(os.path.join(HERE, "golden", "c2_hci_sched.json")))
            Hh = sq.SparseHamiltonian(sq.ChemSystem(C2_FCIDUMP), device=local)
            log = []
            hu, hd, hw, he = spaces.hci_space(Hh, sq.ChemSystem(C2_FCIDUMP), 10**9, eps_schedule=gold["eps_var_sched"], log=log)
            assert [l["n_dets"] for l in log] == gold["n_det"] and [l["nnz_upper"] for l in log] == gold["nnz"]
            assert np.max(np.abs(np.array([l["energy"] for l in log]) - np.array(gold["iter_energy"]))) < 1e-8
            order = np.lexsort((hd[:, 0], hu[:, 0]))
            dig = hashlib.sha256(np.ascontiguousarray(np.stack([hu[order, 0], hd[order, 0]], axis=1)).tobytes()).hexdigest()
            assert dig == gold["sha256_sorted_up_dn_u64"]
            Hh.close()
        # ---- the caller's distribution: determinants owned by hash (get_det_owner), slices in / slices out, against the
        # oracle's emulation of fast_sparse_matrix_multiply_local_band + MPI_REDUCE_SCATTER (do_walk.f90:2259-2260)
        owner = O.det_owner(r["up"], r["dn"], world)
        mine_rows = np.nonzero(owner == rank)[0]
        assert H.set_ownership(owner) == len(mine_rows)
        slices = [x[owner == c] for c in range(world)]
        y_loc = H.matvec_local(slices[rank])
        y_ref_slices = O.matvec_local_band_redscatt(cnt, idx, val, owner, slices)
        assert y_loc.shape == y_ref_slices[rank].shape
        assert np.max(np.abs(y_loc - y_ref_slices[rank])) <= 1e-12 * np.max(np.abs(yref))
        v0 = np.zeros(n)
        v0[0] = 1.0
        gl = H.davidson_sparse_local(n_states=1, initial_vector_local=v0[mine_rows])
        assert gl["ritz"].shape == ref["ritz"].shape and np.max(np.abs(gl["ritz"] - ref["ritz"])) < 1e-8
        assert np.max(np.abs(gl["evecs"][:, 0] - got["evecs"][mine_rows, 0])) < 1e-12     # same vector, only the owned entries
        gl2 = H.davidson_sparse_local(n_states=2)
        assert gl2["n_matvec"] == ref2["n_matvec"] and np.max(np.abs(gl2["evals"] - ref2["evals"])) < 1e-8
        mode = H.exchange_mode()
        tau, e_trial = 0.01, float(val[0])
        H.scale_values(-tau)
        w = x / np.linalg.norm(x)
        dw = H.projector_step(tau, e_trial, w)
        _, dwr = O.projector_step(cnt, idx, -tau * val, tau, e_trial, w)
        assert np.max(np.abs(dw - dwr)) <= 1e-12 * np.max(np.abs(dwr)) + 1e-15
        # 20 projector steps in the distributed form: w_loc += deltaw_loc (do_walk.f90:2321-2323), against the oracle's walk
        w_loc = w[mine_rows].copy()
        H.register_host(w_loc)
        w_ref = w.copy()
        for _ in range(20):
            w_loc += H.projector_step_local(tau, e_trial, w_loc)
            w_ref, _ = O.projector_step(cnt, idx, -tau * val, tau, e_trial, w_ref)
        H.unregister_host(w_loc)
        assert np.max(np.abs(w_loc - w_ref[mine_rows])) <= 1e-11 * np.max(np.abs(w_ref))
        if rank == 0:
            print("multi-gpu parity ok: world=%d time_sym=%s n=%d nnz=%d E=%.10f exchange=%s (local-slice matvec / davidson / projector included)"
                  % (world, time_sym, n, nnz, got["evals"][0], mode))
        H.close()
    # deterministic second-order PT, determinants dealt round-robin to the ranks: the reference's golden HEG numbers
    Sh = O.System.heg(3, 0.5, 14, 7, 1.49)
    rh = Sh.hci(1e-3, n_states=1)
    Hh = sq.SparseHamiltonian(sq.HegSystem(3, 0.5, 14, 7, 1.49), device=local)
    de, nconn = Hh.second_order_pt(rh["up"], rh["dn"], rh["wts"][:, 0], rh["energy"][0], 2e-7)
    assert nconn == 501881 and abs(de - (-0.000939196)) < 5e-10, (nconn, de)     # src/e2e_tests/heg/o_det_ref:431
    if rank == 0:
        print("multi-gpu PT ok: world=%d ndets_connected=%d delta_e=%.9f" % (world, nconn, de))
    # stochastic PT (second_order_pt_alias): every rank draws the same samples, the sampled determinants are dealt round-robin,
    # the partial sums all-reduced -> the first 12 samples of the reference's own log src/e2e_tests/heg/o_st_ref:442-475
    import json
    from conftest import label_sorted
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "heg_o_det_ref.json")))["pt_stochastic"]
    su, sd, sw = label_sorted(rh)
    seed = list(g["irand_seed_1"]); seed[3] = 2 * (seed[3] // 2) + 1
    res = Hh.second_order_pt_alias(su, sd, sw, rh["energy"][0], g["eps_pt"], g["eps_pt_big"], g["n_mc"], g["target_error"], seed, max_samples=12)
    gold = np.array([x["e_now"] for x in g["samples"][:12]])
    assert res["n_samples"] == 12 and np.max(np.abs(res["e_2pt_samples"] - gold)) < 5.1e-10, res["e_2pt_samples"]
    if rank == 0:
        print("multi-gpu stochastic PT ok: world=%d, 12 samples of o_st_ref reproduced, running estimate %.9f" % (world, res["pt_energy"]))
    Hh.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
