set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_9.log)
tail -12 gpurun_out/r02_pytest_gpu_9.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_9.log || exit 1
SQMC_BUILD_PROFILE=1 timeout 1200 python scripts/bundle_inproc.py 10000000 hci "4:14,4:34,4:35,4:14,4:34,2:34,2:14,4:35" > gpurun_out/r02_bundle_ab3.log 2> gpurun_out/r02_bundle_ab3.err
echo "ab3 rc=$?"
cut -c1-200 gpurun_out/r02_bundle_ab3.log
tail -3 gpurun_out/r02_bundle_ab3.err | cut -c1-300
grep "sqmc build" gpurun_out/r02_bundle_ab3.err | tail -12
