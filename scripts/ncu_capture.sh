#!/bin/bash
# ncu evidence for the bench workload (run under gpurun, one GPU).  Each ncu pass runs only after the same command
# exited 0 without ncu.  Outputs land in gpurun_out/ (launch list csv, full-set report of the dominant H.v kernel).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
export SQMC_BENCH_PROFILE_RANGE=1
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
tail -1 gpurun_out/ncu_plain.log | cut -c1-300
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
# one step's worth of H.v kernels (all degree bins) after the three warm-up steps, full metric set
NPER=$(tail -1 gpurun_out/ncu_plain.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(max(1, d['gpu_launches']//d['steps']))")
echo "H.v launches per step: $NPER"
ncu --profile-from-start off --set full --clock-control none --import-source on \
    --kernel-name regex:spmv --launch-skip $((3*NPER)) --launch-count $NPER \
    -o gpurun_out/spmv_full -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/spmv_full.ncu-rep
