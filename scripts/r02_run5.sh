set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q -k "incremental" > gpurun_out/r02_pytest_gpu_5a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_5a.log)
tail -25 gpurun_out/r02_pytest_gpu_5a.log
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_5.log)
tail -8 gpurun_out/r02_pytest_gpu_5.log
SQMC_BUILD_PROFILE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1e7_run5.json 2> gpurun_out/r02_bench_1e7_run5.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1e7_run5.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","value")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["build"], d["parity"]["ok"])
for it in d["hci_iterations"]: print(it)
PY
tail -3 gpurun_out/r02_bench_1e7_run5.err
SQMC_INCREMENTAL=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_1e7_run5_noinc.json 2> gpurun_out/r02_bench_1e7_run5_noinc.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1e7_run5_noinc.json").read().strip().splitlines()[-1])
for it in d["hci_iterations"]: print({k: it[k] for k in ("n_dets","build_s","build_device_ms","build_incremental","select_s","davidson_s")})
PY
