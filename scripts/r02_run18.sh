set -x
cd $GRAFT_REPO_ROOT
for c in 27 28 29; do
SQMC_BUILD_CHUNK_LOG2=$c SQMC_BUILD_PROFILE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_run18_$c.json 2> gpurun_out/r02_bench_run18_$c.err
python - $c <<'PY'
import json,sys
d=json.loads([l for l in open("gpurun_out/r02_bench_run18_%s.json"%sys.argv[1]) if l.startswith("{")][-1])
print("chunk 2^%s" % sys.argv[1], d["build"]["phases_ms"])
PY
done
