set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_45.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_45.log)
tail -4 gpurun_out/r02_pytest_gpu_45.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_45.log || exit 1
python -c "import __graft_entry__ as g; g.smoke()"
( time python bench.py ) > gpurun_out/r02_bench_default_run45.json 2> gpurun_out/r02_bench_default_run45.err
echo "bench rc=$?"; tail -4 gpurun_out/r02_bench_default_run45.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_default_run45.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("ms_per_step","value","gpu_launches")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "build", d["build"]["seconds_wall"], d["build"]["alloc_stall_ms"], d["build"]["phases_ms"], "parity", d["parity"]["ok"], d["clocks"], d["cpu_baseline"]["value"])
print([(it["n_dets"], round(it["build_device_ms"]), round(it["build_alloc_stall_ms"])) for it in d["hci_iterations"]])
PY
for cfg in hubbard heg sweep; do
  timeout 900 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/r02c_config_${cfg}_1gpu.jsonl 2> gpurun_out/r02c_config_${cfg}_1gpu.err
  echo "$cfg rc=$?"
  python - $cfg <<'PY'
import json,sys
for ln in open("gpurun_out/r02c_config_%s_1gpu.jsonl"%sys.argv[1]):
    if not ln.startswith("{"): continue
    d=json.loads(ln)
    print(d["config"]["workload"][:40], d["config"]["n_dets"], round(d["ms_per_step"],4), "frac",round(d["roofline"]["frac"],3), "e2e_ms", round(d["e2e"]["ms_per_step"],3), "build", round(d["build"]["seconds_wall"],3), "stall", round(d["build"]["alloc_stall_ms"]), "parity", d.get("parity",{}).get("ok"))
PY
done
