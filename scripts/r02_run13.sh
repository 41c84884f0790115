set -x
cd $GRAFT_REPO_ROOT
for k in 1 2; do
SQMC_ALLOC_TRACE=1 SQMC_BUILD_PROFILE=2 SQMC_DAV_PROFILE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_run13_$k.json 2> gpurun_out/r02_bench_run13_$k.err
python - $k <<'PY'
import json,sys
d=json.loads([l for l in open("gpurun_out/r02_bench_run13_%s.json"%sys.argv[1]) if l.startswith("{")][-1])
print(d["build"]["seconds_wall"], d["build"]["phases_ms"]["total_ms"], d["ms_per_step"], d["build"]["space_seconds"])
print([ (it["n_dets"], round(it["build_device_ms"]), round(it["select_s"]*1e3), round(it["davidson_s"]*1e3)) for it in d["hci_iterations"]])
PY
grep "sqmc alloc\|davidson\]" gpurun_out/r02_bench_run13_$k.err | cut -c1-200 | tail -40
done
