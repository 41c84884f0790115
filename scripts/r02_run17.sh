set -x
cd $GRAFT_REPO_ROOT
export SQMC_BENCH_PROFILE_RANGE=0
CMD="python bench.py --n-dets 4000000 --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/r02_ncu_build_plain.log 2>&1 || { tail -5 gpurun_out/r02_ncu_build_plain.log; exit 1; }
# the last launches of the build kernels belong to the from-scratch build of the final space
for k in connect_bitmap eval_kernel compact_copy bundle_encode; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 60 --launch-count 2 -o gpurun_out/r02_build_$k -f $CMD > gpurun_out/r02_ncu_build_$k.log 2>&1
  echo "$k rc=$?"; ls -la gpurun_out/r02_build_$k.ncu-rep
done
