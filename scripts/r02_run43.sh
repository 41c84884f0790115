set -x
cd $GRAFT_REPO_ROOT
export SQMC_BENCH_PROFILE_RANGE=0
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
# the last launches of the build kernels belong to the from-scratch build of the final 10^7 space (launch counts: profiles/r02c_launches_hci1e7_summary.txt)
ncu --set full --clock-control none --import-source on --kernel-name regex:connect_bitmap --launch-skip 112 --launch-count 2 -o gpurun_out/r02c_build_connect_bitmap -f $CMD > gpurun_out/r02c_ncu_build_connect_bitmap.log 2>&1
echo "connect rc=$?"; ls -la gpurun_out/r02c_build_connect_bitmap.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name regex:eval_kernel --launch-skip 105 --launch-count 2 -o gpurun_out/r02c_build_eval_kernel -f $CMD > gpurun_out/r02c_ncu_build_eval_kernel.log 2>&1
echo "eval rc=$?"; ls -la gpurun_out/r02c_build_eval_kernel.ncu-rep
