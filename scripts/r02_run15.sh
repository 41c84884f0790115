set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_15.log)
tail -12 gpurun_out/r02_pytest_gpu_15.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_15.log || exit 1
for k in 1 0; do
SQMC_CONNECT_BITMAP=$k SQMC_BUILD_PROFILE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_run15_bmp$k.json 2> gpurun_out/r02_bench_run15_bmp$k.err
python - $k <<'PY'
import json,sys
d=json.loads([l for l in open("gpurun_out/r02_bench_run15_bmp%s.json"%sys.argv[1]) if l.startswith("{")][-1])
print("bitmap", sys.argv[1], d["build"], d["parity"]["ok"])
print([ (it["n_dets"], round(it["build_device_ms"])) for it in d["hci_iterations"]])
PY
grep "sqmc build" gpurun_out/r02_bench_run15_bmp$k.err | tail -11 | cut -c1-110
done
