#!/usr/bin/env python
"""The HCI variational loop entirely on one B200 (selection -> H build -> Davidson through the C ABI), following
perform_hci (hci.f90:359-517) with a decreasing eps_var schedule until the space reaches a target size: the
"C2 cc-pVDZ HCI with eps_var lowered to give ~10^7 determinants" recipe of BASELINE.json configs[3] (SURVEY S4 primary).
Prints one JSON line per HCI iteration."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import sqmc_b200 as sq
    ap = argparse.ArgumentParser()
    ap.add_argument("--r", default="1.24253")
    ap.add_argument("--time-sym", action="store_true")
    ap.add_argument("--eps", default="1e-3,3e-4,1e-4,3e-5,1e-5,5e-6,2e-6,1e-6,5e-7")
    ap.add_argument("--target", type=int, default=10_000_000)
    ap.add_argument("--iters-per-eps", type=int, default=1)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:   # torchrun: selection, build and Davidson are all sharded over the ranks
        import torch
        import torch.distributed as dist
        from sqmc_b200 import _lib
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        obj = [_lib.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        _lib.init(device=local, rank=rank, nranks=world, unique_id=obj[0])
    cs = sq.ChemSystem(os.path.join(ROOT, "data", "C2_v2z_curve", "r" + args.r, "FCIDUMP"), time_sym=args.time_sym, z=1)
    H = sq.SparseHamiltonian(cs, device=local)
    up = sq.dets_to_u64([cs.hf_up])
    dn = sq.dets_to_u64([cs.hf_dn])
    wts = np.ones((1, 1))
    min_h = np.full(1, 9e99)
    it = 0
    t_start = time.perf_counter()
    for eps in [float(x) for x in args.eps.split(",")]:
        for _ in range(args.iters_per_eps):
            it += 1
            n_old = len(up)
            coeffs = np.abs(wts[:, 0])
            t0 = time.perf_counter()
            nu, nd, min_h = H.get_next_det_list(up, dn, coeffs, min_h, eps)
            t_sel = time.perf_counter() - t0
            if len(nu) == 0:
                continue
            if n_old + len(nu) > args.target:          # keep the first `target - n_old` new determinants (sorted by label)
                nu, nd = nu[:args.target - n_old], nd[:args.target - n_old]
            up, dn = np.concatenate([up, nu]), np.concatenate([dn, nd])
            min_h = np.concatenate([min_h, np.full(len(nu), 9e99)])
            n = len(up)
            t0 = time.perf_counter()
            nnz = H.generate_sparse_ham_upper_triangular(up, dn, ndet_old=n_old)
            t_build = time.perf_counter() - t0
            v0 = np.zeros((n, 1))
            v0[:n_old, 0] = wts[:, 0]
            t0 = time.perf_counter()
            d = H.davidson_sparse(n_states=1, initial_vector=v0)
            t_dav = time.perf_counter() - t0
            wts = d["evecs"]
            if rank == 0:
              print(json.dumps({"n_gpus": world, "iter": it, "eps_var": eps, "n_dets": n, "n_new": int(len(nu)), "nnz_upper": int(nnz), "energy": float(d["evals"][0]),
                                "n_matvec": d["n_matvec"], "select_s": t_sel, "build_s": t_build, "davidson_s": t_dav,
                                "elapsed_s": time.perf_counter() - t_start}), flush=True)
            if n >= args.target:
                return


if __name__ == "__main__":
    main()
