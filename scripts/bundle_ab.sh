#!/bin/bash
# A/B of the row-bundle H.v ordering (SQMC_BUNDLE=0|2|4|8): parity tests with bundling on, then bench lines.
set -u
mkdir -p gpurun_out
for R in 8 2; do
  SQMC_BUNDLE=$R timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size and not wcsr and not multi" > gpurun_out/bundle_pytest_$R.log 2>&1
  echo "pytest SQMC_BUNDLE=$R rc=$?"; tail -3 gpurun_out/bundle_pytest_$R.log
done
for R in 0 2 4 8; do
  SQMC_BUNDLE=$R timeout 600 python bench.py --space lowest --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bundle_bench_$R.log 2>&1
  echo "bench SQMC_BUNDLE=$R rc=$?"
  tail -1 gpurun_out/bundle_bench_$R.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['frac'], d['build']['seconds_wall'], d['build'].get('ms'), d['clocks'])"
done
