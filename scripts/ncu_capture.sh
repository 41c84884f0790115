#!/bin/bash
# ncu evidence for the bench workload (run under gpurun, one GPU).  Each ncu pass runs only after the same command
# exited 0 without ncu.  Outputs land in gpurun_out/ (launch lists, full-set report of the H.v kernel).
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
export SQMC_BENCH_PROFILE_RANGE=1
$CMD > gpurun_out/${TAG}_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_ncu_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_ncu_plain.log | cut -c1-300
# (1) launch list of the warm-up + timed steps only (cudaProfilerStart/Stop around them)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_timed_region.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "timed-region launch list rc=$?"; wc -l gpurun_out/${TAG}_launches_timed_region.csv
# (2) one H.v launch after the warm-up steps, full metric set, with source
ncu --profile-from-start off --set full --clock-control none --import-source on \
    --kernel-name regex:bundle_hv --launch-skip 3 --launch-count 1 \
    -o gpurun_out/${TAG}_bundle_hv_full -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/${TAG}_bundle_hv_full.ncu-rep
# (3) whole-program launch list (HCI loop that grows the space + final build + H.v): which kernels make up the program
unset SQMC_BENCH_PROFILE_RANGE
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_program.csv $CMD > gpurun_out/${TAG}_ncu_program.log 2>&1
echo "program launch list rc=$?"; wc -l gpurun_out/${TAG}_launches_program.csv; gzip -f gpurun_out/${TAG}_launches_program.csv
