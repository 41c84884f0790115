#!/usr/bin/env python
"""Throughput of the hot path on the BASELINE.json configurations that are not the bench workload
(configs 2, 3 and 5): H build (stored upper nnz/s) and H.v / projector step (nnz_full/s, algorithmic GB/s).
Single process (1 GPU) or torchrun (row-sharded).  Prints one JSON object per case; results are copied to
profiles/ by hand.  Timing: wall clock around synchronised calls for the build, CUDA events on the launching
stream for H.v (device-resident vectors)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import sqmc_b200 as sq
    from sqmc_b200 import _lib, spaces
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="hubbard,heg,sweep")
    ap.add_argument("--sweep-dets", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        obj = [_lib.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        uid = obj[0]
    _lib.init(device=local, rank=rank, nranks=world, unique_id=uid)
    L = _lib.load()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def run_case(name, system, up, dn, extra=None):
        H = sq.SparseHamiltonian(system, device=local)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nnz_upper = H.generate_sparse_ham_upper_triangular(up, dn)
        torch.cuda.synchronize()
        t_build = time.perf_counter() - t0
        info = H.nnz()
        n, nnz_full = info["n"], info["nnz_full"]
        nloc, _ = H.local_rows()
        x = torch.from_numpy(spaces.splitmix_vector(n)).cuda()
        y = torch.zeros(max(nloc, 1), dtype=torch.float64, device="cuda")
        sp = C.c_void_p(stream.cuda_stream)
        for _ in range(3):
            _lib.check(L.sqmc_b200_matvec_dev(H._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), sp))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            _lib.check(L.sqmc_b200_matvec_dev(H._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), sp))
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms, t_build], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, t_build = float(t[0]), float(t[1])
        out = {"case": name, "n_gpus": world, "n_dets": n, "nnz_upper": nnz_upper, "nnz_full": nnz_full, "build_s": t_build,
               "build_nnz_upper_per_s": nnz_upper / t_build, "hv_ms": ms, "hv_nnz_per_s": nnz_full / (ms * 1e-3),
               "hv_algorithmic_GBs": (12.0 * nnz_full + 20.0 * n) / (ms * 1e-3) / 1e9}
        if extra:
            out.update(extra(H))
        if rank == 0:
            print(json.dumps(out), flush=True)
        H.close()

    cases = args.cases.split(",")
    if "hubbard" in cases:  # config 2: 4x4 half filling, momentum sector (0,0), deterministic-space projector
        hub = sq.HubbardKSystem(4, 4, 1.0, 4.0, 8, 8)
        for N in (100_000, 1_000_000, None):
            up, dn, total = spaces.hubbard_momentum_sector(hub, N)

            def proj(H, up=up, dn=dn):
                n = len(up)
                w = np.zeros(n)
                w[0] = 1.0
                tau = 0.01
                H.scale_values(-tau)
                H.projector_step(tau, -10.0, w)
                t0 = time.perf_counter()
                for _ in range(5):
                    w = w + H.projector_step(tau, -10.0, w)
                return {"projector_step_host_vectors_ms": (time.perf_counter() - t0) / 5 * 1e3}
            run_case("hubbard4x4_half_filling_k0_N=%s" % (N if N else "full(%d)" % total), hub, up, dn, proj if world == 1 else None)
    if "heg" in cases:  # config 3: HEG 14 electrons, cutoff 2.0 (33 orbitals), heat-bath space from the oracle (input generator)
        from oracle import oracle as O
        S = O.System.heg(3, 0.5, 14, 7, 2.0)
        r = S.hci(2e-4, n_states=1, max_iters=2)
        run_case("heg14_rs0.5_cutoff2.0_eps2e-4", sq.HegSystem(3, 0.5, 14, 7, 2.0), r["up"], r["dn"])
    if "sweep" in cases:  # config 5: C2 binding curve, 9 geometries
        for rr in ("1.0", "1.1", "1.2", "1.24253", "1.3", "1.4", "1.6", "1.8", "2.0"):
            chem = sq.ChemSystem(os.path.join(ROOT, "data", "C2_v2z_curve", "r" + rr, "FCIDUMP"))
            up, dn, _ = spaces.c2_lowest_energy_space(chem, args.sweep_dets)
            run_case("c2_ccpvdz_r%s_N=%d" % (rr, args.sweep_dets), chem, up, dn)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
