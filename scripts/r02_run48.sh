set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_launches_program.csv $CMD > gpurun_out/r02d_ncu_program.log 2>&1
echo "program launch list rc=$?"; wc -l gpurun_out/r02d_launches_program.csv; gzip -f gpurun_out/r02d_launches_program.csv
