#!/usr/bin/env python
"""bench.py -- sparse H.v throughput (nnz/s, achieved HBM GB/s) on B200, with the CPU baseline beside it.

Workload (BASELINE.json configs[3]): C2 cc-pVDZ r=1.24253 (data/C2_v2z_curve), time_sym=f,
the N lowest-diagonal-energy A_g determinants (default N = 10^7), H built on
the GPU by this library, then K timed H.v products (the Davidson matvec, more_tools.f90:2188).
A "step" is one H.v over the whole matrix.  Under torchrun (N GPUs) rows of H are sharded,
the vector is all-gathered over NCCL every step (strong scaling: the matrix is fixed).

  value : nnz_full / t  with x, y and H resident in HBM (CUDA events on the launching stream)
  e2e   : same through sqmc_b200_matvec with HOST vectors (H2D + D2H inside the timed region)
  roofline.achieved : algorithmic bytes (12*nnz_full + 20*n) / event time

--impl reference times the CPU restatement of the reference's own
fast_sparse_matrix_multiply_upper_triangular (oracle/, threads emulate the MPI rank
decomposition of davidson_sparse_mpi2) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FCIDUMP = os.path.join(ROOT, "data", "C2_v2z_curve", "r1.24253", "FCIDUMP")
METRIC = "sparse_Hv_nnz_per_s"
UNIT = "nnz/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [names[k] for k in range(4) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


CPU_SAMPLE_SCHED = (1e-3, 3e-4, 1e-4)


def cpu_sample(args):
    """The bounded CPU sample of the workload, built with the oracle only: for --space hci the same HCI run stopped at
    eps_var = 1e-4 (171,060 determinants; tests/golden/c2_hci_sched.json), for --space lowest the cpu_sample_dets
    lowest-energy determinants.  -> (counts, idx, val, description, seconds to obtain the sample matrix)"""
    import sqmc_b200 as sq
    from sqmc_b200 import spaces
    from oracle import oracle as O
    chem = sq.ChemSystem(FCIDUMP)
    S = O.System.chem(FCIDUMP, chem.norb, chem.nelec, chem.nup, chem.orbital_symmetries_fcidump)
    t0 = time.perf_counter()
    if args.space == "hci":
        r = S.hci(CPU_SAMPLE_SCHED[-1], eps_var_sched=CPU_SAMPLE_SCHED, max_iters=len(CPU_SAMPLE_SCHED))
        up, dn = r["up"], r["dn"]
        desc = "the same HCI run stopped at eps_var=1e-4 (oracle perform_hci, %d dets)" % len(up)
    else:
        up, dn, _ = spaces.c2_lowest_energy_space(chem, args.cpu_sample_dets)
        desc = "the %d lowest-energy A_g dets" % len(up)
    t_space = time.perf_counter() - t0
    t0 = time.perf_counter()
    cnt, idx, val = S.build_upper(up, dn)
    t_build = time.perf_counter() - t0
    return cnt, idx, val, desc, t_space, t_build


def workload_name(args, n, space_desc=None):
    if args.space == "hci":
        d = space_desc or "HCI space (heat-bath selection + build + Davidson per iteration, eps_var lowered 1e-3 -> 5e-7)"
    else:
        d = space_desc or "lowest-diagonal-energy A_g determinants"
    return "C2 cc-pVDZ r1.24253 (FCIDUMP), time_sym=f, %d determinants: %s; step = one H.v" % (n, d)


def run_reference(args):
    """CPU arm: the oracle's restatement of fast_sparse_matrix_multiply_upper_triangular on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sqmc_b200 import spaces
    from oracle import oracle as O
    cnt, idx, val, desc, t_space, t_build = cpu_sample(args)
    n = len(cnt)
    nnz_full = 2 * len(idx) - n
    cores = os.cpu_count() or 1
    x = spaces.splitmix_vector(n)
    for _ in range(args.warmup):
        O.matvec_upper_mt(cnt, idx, val, x, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.matvec_upper_mt(cnt, idx, val, x, cores)
    t = (time.perf_counter() - t0) / args.steps
    v = nnz_full / t
    sample = "C2 cc-pVDZ r1.24253 time_sym=f, %s (nnz_full=%d); %d threads, rows dealt round-robin, private y + reduction" % (desc, nnz_full, cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, args.n_dets), "space": args.space, "n_dets": args.n_dets,
                       "sample_n_dets": n, "sample_nnz_full": nnz_full,
                       "note": "each step is one CPU H.v on the bounded sample described in cpu_baseline.sample"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "build_nnz_upper_per_s_1core": len(idx) / t_build},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-dets", type=int, default=int(os.environ.get("SQMC_BENCH_NDETS", 10_000_000)))
    ap.add_argument("--space", default=os.environ.get("SQMC_BENCH_SPACE", "hci"), choices=["hci", "lowest"],
                    help="hci (BASELINE.json configs[3] recipe): an HCI run on the GPU (selection + build + Davidson through the "
                         "library) with eps_var lowered until N determinants; lowest: the N lowest-diagonal-energy A_g "
                         "determinants (host generated, SURVEY.md S4 fallback)")
    ap.add_argument("--cpu-sample-dets", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--davidson", action="store_true", help="also run a full device Davidson and report it in extra")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import ctypes as C
    import sqmc_b200 as sq
    from sqmc_b200 import _lib, spaces

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        obj = [_lib.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        uid = obj[0]
    _lib.init(device=local_rank, rank=rank, nranks=world, unique_id=uid)
    L = _lib.load()

    # ---- workload (host, untimed)
    t0 = time.perf_counter()
    chem = sq.ChemSystem(FCIDUMP)
    H = sq.SparseHamiltonian(chem, device=local_rank)
    hci_log = []
    if args.space == "hci":
        up, dn, _, e_var = spaces.hci_space(H, chem, args.n_dets, log=hci_log)
        sector = 27944940
        space_desc = "HCI space grown on the GPU (heat-bath selection + build + Davidson), eps_var lowered to %.0e, E_var = %.9f Ha" % (
            hci_log[-1]["eps_var"], e_var)
    else:
        up, dn, sector = spaces.c2_lowest_energy_space(chem, args.n_dets)
        space_desc = "%d lowest-diagonal-energy A_g determinants of %d" % (len(up), sector)
    n = len(up)
    t_space = time.perf_counter() - t0
    t0 = time.perf_counter()
    nnz_upper = H.generate_sparse_ham_upper_triangular(up, dn)   # timed from-scratch build of the final space
    t_build = time.perf_counter() - t0
    info = H.nnz()
    nnz_full = info["nnz_full"]
    nloc, nnz_loc = H.local_rows()
    bt = H.build_times()
    if world > 1:
        allnnz = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(allnnz, torch.tensor([nnz_loc], dtype=torch.int64, device="cuda"))
        nnz_per_rank = [int(t.item()) for t in allnnz]
    else:
        nnz_per_rank = [nnz_loc]
    launches_before = L.sqmc_b200_launch_count()

    # ---- resident vectors (internal row order: the order Davidson keeps its Krylov vectors in)
    stream = torch.cuda.Stream()  # a real (non-default) stream: the library launches on the stream it is handed
    torch.cuda.set_stream(stream)
    x_host = spaces.splitmix_vector(n)
    x = torch.from_numpy(x_host).cuda()
    y = torch.zeros(max(nloc, 1), dtype=torch.float64, device="cuda")
    sptr = C.c_void_p(stream.cuda_stream)

    def step_dev():
        _lib.check(L.sqmc_b200_matvec_dev(H._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), sptr))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    prof_range = os.environ.get("SQMC_BENCH_PROFILE_RANGE") == "1"   # ncu --profile-from-start off: only warm-up + timed steps
    if prof_range:
        torch.cuda.profiler.start()
    for _ in range(args.warmup):
        step_dev()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    l0 = L.sqmc_b200_launch_count()
    ev[0].record(stream)
    for k in range(args.steps):
        step_dev()
        ev[k + 1].record(stream)
    barrier()
    l1 = L.sqmc_b200_launch_count()
    if prof_range:
        torch.cuda.profiler.stop()
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    ms_total = ev[0].elapsed_time(ev[args.steps])
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host vectors through the C ABI (pinned host memory, H2D + D2H inside)
    xh = torch.from_numpy(x_host.copy()).pin_memory()
    yh = torch.zeros(n, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def step_e2e():
        _lib.check(L.sqmc_b200_matvec(H._h, C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()), 1, n))

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())

    extra = {}
    if args.davidson:
        t0 = time.perf_counter()
        dv = H.davidson_sparse(n_states=1)
        extra["davidson"] = {"seconds": time.perf_counter() - t0, "n_matvec": dv["n_matvec"], "energy": float(dv["evals"][0])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the H.v launch group (bins of one CSR; the warp-per-row kernel dominates)
    peak, peak_src = measured_peak_gbs()
    # per-rank algorithmic bytes (the slowest rank bounds the step; rows are nnz-balanced)
    alg_bytes = 12.0 * nnz_loc + 20.0 * nloc if world > 1 else 12.0 * nnz_full + 20.0 * n
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "spmv_dram_bytes_per_launch.json")
    if os.path.exists(tp) and world == 1:  # only when the committed ncu capture is of this exact matrix
        try:
            tj = json.load(open(tp))
            if int(tj.get("nnz_full", -1)) == int(nnz_full):
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": nnz_full / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, n), "space": args.space, "space_detail": space_desc, "n_dets": n, "nnz_upper": nnz_upper, "nnz_full": nnz_full, "parallelism": "rows x%d" % world,
                   "l2": "inputs larger than L2 (matrix %.1f GB streamed per step)" % (12.0 * nnz_full / 1e9)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": alg_bytes, "frac_of_nominal_8TBs": achieved / 8000.0},
        "e2e": {"value": nnz_full / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n, "ms_per_step": t_e2e * 1e3},
        "gpu_launches": int(l1 - l0),
        "clocks": clocks,
        "build": {"seconds_wall": t_build, "nnz_upper_per_s": nnz_upper / t_build, "phases_ms": bt, "space_seconds": t_space},
        "step_ms_min_max": [min(per), max(per)],
        "nnz_per_rank_max_over_mean": max(nnz_per_rank) / (sum(nnz_per_rank) / len(nnz_per_rank)),
    }
    if hci_log:
        line["hci_iterations"] = hci_log
    line.update(extra)
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    """oracle (port of the reference algorithm) on a bounded sample, 1 core: ~10-30 s of CPU work."""
    from sqmc_b200 import spaces
    from oracle import oracle as O
    cnt, idx, val, desc, t_space, t_build = cpu_sample(args)
    n = len(cnt)
    nnz_full = 2 * len(idx) - n
    x = spaces.splitmix_vector(n)
    O.matvec_upper(cnt, idx, val, x)
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 5.0:
        O.matvec_upper(cnt, idx, val, x)
        reps += 1
    t = (time.perf_counter() - t0) / reps
    return {"value": nnz_full / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle fast_sparse_matrix_multiply_upper_triangular on %s of the same C2 workload (nnz_full=%d), %d reps; "
                      "obtaining the sample's space took the oracle %.1f s, its from-scratch H build %.1f s" % (desc, nnz_full, reps, t_space, t_build),
            "build_nnz_upper_per_s": len(idx) / t_build}


if __name__ == "__main__":
    main()
