set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_7.log)
tail -15 gpurun_out/r02_pytest_gpu_7.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_7.log || exit 1
SQMC_BUILD_PROFILE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1e7_run7.json 2> gpurun_out/r02_bench_1e7_run7.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1e7_run7.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","value")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["build"], d["parity"]["ok"])
for it in d["hci_iterations"]: print({k: it[k] for k in ("n_dets","build_s","build_device_ms","build_incremental","select_s","davidson_s")})
PY
grep "sqmc build" gpurun_out/r02_bench_1e7_run7.err | awk 'BEGIN{b=0} /free previous/{b++} {print b": "$0}' | awk -F: '$1>=8' | cut -c1-120
