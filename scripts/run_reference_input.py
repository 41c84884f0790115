#!/usr/bin/env python
"""Run one of the reference's own `hci` input files on the GPU path (system set-up, variational HCI loop, deterministic PT):
  python scripts/run_reference_input.py data/C2_v2z_curve/r1.24253/i_1sigma_g
  python scripts/run_reference_input.py tests/golden/heg_i_det          (src/e2e_tests/heg/i_det; golden output o_det_ref)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqmc_b200 import hci

t0 = time.perf_counter()
res = hci.run_input(sys.argv[1], fcidump=sys.argv[2] if len(sys.argv) > 2 else None)
print(json.dumps({"input": sys.argv[1], "seconds": time.perf_counter() - t0, "n_det": len(res["up"]), "energies": [float(e) for e in res["energy"]],
                  "pt": [{"delta_e": d, "ndets_connected": n} for d, n in res["pt"]],
                  "total": [float(e) + d for e, (d, n) in zip(res["energy"], res["pt"])]}))
