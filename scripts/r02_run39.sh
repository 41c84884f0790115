set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_39.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_39.log)
tail -4 gpurun_out/r02_pytest_gpu_39.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_39.log || exit 1
python -c "import __graft_entry__ as g; g.smoke()"
( time python bench.py ) > gpurun_out/r02_bench_default_run39.json 2> gpurun_out/r02_bench_default_run39.err
echo "bench rc=$?"; tail -4 gpurun_out/r02_bench_default_run39.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_default_run39.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("ms_per_step","value","gpu_launches")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "build", d["build"]["seconds_wall"], d["build"]["phases_ms"], "parity", d["parity"]["ok"], d["clocks"], d["cpu_baseline"]["value"])
for it in d["hci_iterations"]: print({k: it[k] for k in ("n_dets","build_s","build_device_ms","build_incremental","select_s","davidson_s")})
PY
