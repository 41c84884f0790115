set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_46.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_46.log)
tail -3 gpurun_out/r02_pytest_gpu_46.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_46.log || exit 1
for cfg in heg; do
  timeout 900 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/r02c_config_${cfg}_1gpu.jsonl 2> gpurun_out/r02c_config_${cfg}_1gpu.err
  echo "$cfg rc=$?"
  python - $cfg <<'PY'
import json,sys
for ln in open("gpurun_out/r02c_config_%s_1gpu.jsonl"%sys.argv[1]):
    if not ln.startswith("{"): continue
    d=json.loads(ln)
    print(d["config"]["workload"][:40], d["config"]["n_dets"], round(d["ms_per_step"],4), "frac",round(d["roofline"]["frac"],3), "e2e_ms", round(d["e2e"]["ms_per_step"],3), "build", round(d["build"]["seconds_wall"],3), {k:round(v,1) for k,v in d["build"]["phases_ms"].items() if k.endswith("ms")}, "parity", d.get("parity",{}).get("ok"))
PY
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_run46.json 2> gpurun_out/r02_bench_run46.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_run46.json") if l.startswith("{")][-1])
print(round(d["ms_per_step"],3), "build", round(d["build"]["seconds_wall"],3), round(d["build"]["alloc_stall_ms"]), {k:round(v) for k,v in d["build"]["phases_ms"].items() if k.endswith("ms")}, "parity", d["parity"]["ok"])
print([(it["n_dets"], round(it["build_device_ms"]), round(it["build_alloc_stall_ms"])) for it in d["hci_iterations"]])
PY
