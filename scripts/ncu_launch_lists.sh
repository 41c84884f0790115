#!/bin/bash
# whole-program launch lists (build + H.v for the host-generated space; the full HCI loop for the default space).
set -u
mkdir -p gpurun_out
A="python bench.py --space lowest --steps 3 --warmup 3 --no-cpu-baseline"
B="python bench.py --space hci --steps 3 --warmup 3 --no-cpu-baseline"
if [ "${ONLY_HCI:-0}" != "1" ]; then
$A > gpurun_out/ll_plain_lowest.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_lowest.csv $A > gpurun_out/ll_ncu_lowest.log 2>&1
echo "lowest rc=$?"; wc -l gpurun_out/launches_lowest.csv
fi
$B > gpurun_out/ll_plain_hci.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file gpurun_out/launches_hci.csv $B > gpurun_out/ll_ncu_hci.log 2>&1
echo "hci rc=$?"; wc -l gpurun_out/launches_hci.csv
gzip -f gpurun_out/launches_hci.csv
ls -la gpurun_out/
