#!/bin/bash
set -u
mkdir -p gpurun_out
export SQMC_BENCH_PROFILE_RANGE=1
for R in ${RS:-2 4 8}; do
  export SQMC_BUNDLE=$R
  CMD="python bench.py --space lowest --steps 3 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/bncu_plain_$R.log 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name regex:spmv --launch-skip 3 --launch-count 1 \
      -o gpurun_out/bundle_full_$R -f $CMD > gpurun_out/bncu_full_$R.log 2>&1
  echo "R=$R rc=$?"
done
ls -la gpurun_out/*.ncu-rep
