set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_incremental.py tests/test_gpu_edge_cases.py tests/test_gpu_configs.py tests/test_gpu_bundle.py -x -q > gpurun_out/r02_pytest_gpu_40.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_40.log)
tail -4 gpurun_out/r02_pytest_gpu_40.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_40.log || exit 1
for k in 1 2; do
SQMC_ALLOC_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_run40_$k.json 2> gpurun_out/r02_bench_run40_$k.err
echo "bench rc=$?"
python - $k <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/r02_bench_run40_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print({k:d[k] for k in ("ms_per_step",)}, "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["ms_per_step"],2), "build", round(d["build"]["seconds_wall"],3), "stall", round(d["build"]["alloc_stall_ms"]), {k:round(v) for k,v in d["build"]["phases_ms"].items() if k.endswith("ms")}, "parity", d["parity"]["ok"])
print([(it["n_dets"], round(it["build_device_ms"]), round(it["build_alloc_stall_ms"])) for it in d["hci_iterations"]])
PY
grep "sqmc alloc" gpurun_out/r02_bench_run40_$k.err | sort -t' ' -k12 -n | tail -5
done
