set -x
cd $GRAFT_REPO_ROOT
SQMC_BUILD_PROFILE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1e7_run8.json 2> gpurun_out/r02_bench_1e7_run8.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1e7_run8.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","value")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["build"], d["parity"]["ok"])
for it in d["hci_iterations"]: print({k: it[k] for k in ("n_dets","build_s","build_device_ms","build_incremental","select_s","davidson_s")})
PY
grep "sqmc build" gpurun_out/r02_bench_1e7_run8.err | awk 'BEGIN{b=0} /free previous/{b++} {print b": "$0}' | awk -F: '$1>=8' | cut -c1-120
for cfg in hubbard heg sweep; do
  timeout 1200 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/r02_config_${cfg}_1gpu.jsonl 2> gpurun_out/r02_config_${cfg}_1gpu.err
  echo "$cfg rc=$?"; tail -3 gpurun_out/r02_config_${cfg}_1gpu.err
  python - $cfg <<'PY'
import json,sys
for ln in open("gpurun_out/r02_config_%s_1gpu.jsonl"%sys.argv[1]):
    ln=ln.strip()
    if not ln.startswith("{"): continue
    d=json.loads(ln)
    print(d["config"]["workload"][:70], d["config"]["n_dets"], "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), "e2e_ms", round(d["e2e"]["ms_per_step"],3), "build_s", round(d["build"]["seconds_wall"],3), "nnz_up/s", "%.3g"%d["build"]["nnz_upper_per_s"], "parity", d.get("parity",{}).get("ok"))
PY
done
bash scripts/ncu_capture.sh r02
timeout 1200 python scripts/bundle_inproc.py 10000000 hci "4:14,4:24,4:26,4:28,4:15,2:14,2:24,2:26,4:14,4:26" > gpurun_out/r02_bundle_ab2.log 2> gpurun_out/r02_bundle_ab2.err
echo "ab2 rc=$?"
cut -c1-220 gpurun_out/r02_bundle_ab2.log
tail -3 gpurun_out/r02_bundle_ab2.err
