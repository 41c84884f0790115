set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q -k "incremental or parity or edge or configs or select" > gpurun_out/r02_pytest_gpu_16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_16.log)
tail -6 gpurun_out/r02_pytest_gpu_16.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_16.log || exit 1
SQMC_BUILD_PROFILE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_run16.json 2> gpurun_out/r02_bench_run16.err
python - <<'PY'
import json,sys
d=json.loads([l for l in open("gpurun_out/r02_bench_run16.json") if l.startswith("{")][-1])
print(d["build"], d["parity"]["ok"])
print([ (it["n_dets"], round(it["build_device_ms"])) for it in d["hci_iterations"]])
PY
grep "sqmc build" gpurun_out/r02_bench_run16.err | tail -22 | cut -c1-110
