#!/bin/bash
set -u
mkdir -p gpurun_out
for R in ${RS:-0 2 4 8}; do
  SQMC_BUNDLE=$R timeout 600 python bench.py --space lowest --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bundle_bench_$R.log 2>&1
  echo "bench SQMC_BUNDLE=$R rc=$?"
  tail -1 gpurun_out/bundle_bench_$R.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['frac'], d['build']['seconds_wall'], d['clocks'])"
done
