set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_28.log)
tail -6 gpurun_out/r02_pytest_gpu_28.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_28.log || exit 1
( time python bench.py ) > gpurun_out/r02_bench_default_run28.json 2> gpurun_out/r02_bench_default_run28.err
echo "bench rc=$?"; tail -4 gpurun_out/r02_bench_default_run28.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_default_run28.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("ms_per_step","value","gpu_launches")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "build", d["build"]["seconds_wall"], "parity", d["parity"]["ok"], d["clocks"], d["cpu_baseline"]["value"])
PY
timeout 600 python scripts/bundle_inproc.py 10000000 hci 4:44,4:34,4:44,4:34 > gpurun_out/r02_bundle_ab11.log 2>&1
grep '^{' gpurun_out/r02_bundle_ab11.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 3), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
