#!/usr/bin/env python
"""HCI + deterministic second-order PT entirely on the GPU for C2 cc-pVDZ (time_sym = f): variational stage with a decreasing
eps_var schedule, then sqmc_b200_pt2 for a list of eps_pt.  Prints one JSON line per PT call (wall seconds, connections)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    import sqmc_b200 as sq
    from sqmc_b200 import spaces
    ap = argparse.ArgumentParser()
    ap.add_argument("--eps-var", default="1e-3,3e-4,1e-4")
    ap.add_argument("--eps-pt", default="1e-5,1e-6,1e-7")
    ap.add_argument("--oracle", action="store_true", help="also run the CPU oracle (slow: enumerates every double excitation)")
    args = ap.parse_args()
    cs = sq.ChemSystem(os.path.join(ROOT, "data", "C2_v2z_curve", "r1.24253", "FCIDUMP"))
    H = sq.SparseHamiltonian(cs)
    t0 = time.perf_counter()
    up, dn, wts, e = spaces.hci_space(H, cs, 10**9, eps_schedule=[float(x) for x in args.eps_var.split(",")])
    t_var = time.perf_counter() - t0
    H.second_order_pt(up[:100], dn[:100], wts[:100, 0], e, 1e-3)   # builds the heat-bath tables, warms the kernels
    for eps_pt in [float(x) for x in args.eps_pt.split(",")]:
        t0 = time.perf_counter()
        de, nc = H.second_order_pt(up, dn, wts[:, 0], e, eps_pt)
        t = time.perf_counter() - t0
        rec = {"n_dets": len(up), "variational_seconds": t_var, "E_var": e, "eps_pt": eps_pt, "delta_e_2pt": de, "E_total": e + de,
               "ndets_connected": nc, "pt_seconds": t}
        if args.oracle:
            from oracle import oracle as O
            S = O.System.chem(os.path.join(ROOT, "data", "C2_v2z_curve", "r1.24253", "FCIDUMP"), cs.norb, cs.nelec, cs.nup, cs.orbital_symmetries_fcidump)
            t0 = time.perf_counter()
            ode, onc = S.pt2(up, dn, wts[:, 0], e, eps_pt)
            rec.update({"oracle_seconds_1core": time.perf_counter() - t0, "oracle_delta_e_2pt": ode, "oracle_ndets_connected": onc})
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
