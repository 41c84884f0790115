set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q -k "incremental" > gpurun_out/r02_pytest_gpu_11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_11.log)
tail -5 gpurun_out/r02_pytest_gpu_11.log
SQMC_BUILD_PROFILE=2 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_run11.json 2> gpurun_out/r02_bench_run11.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_run11.json") if l.startswith("{")][-1])
print(d["build"])
for it in d["hci_iterations"]: print({k: it[k] for k in ("n_dets","build_s","build_device_ms","build_incremental")})
PY
grep "sqmc build" gpurun_out/r02_bench_run11.err | awk 'BEGIN{b=0} /free previous/{b++} {print b": "$0}' | awk -F: '$1>=9' | cut -c1-120
