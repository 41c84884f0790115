set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q -k "bundle or local or parity or edge" > gpurun_out/r02_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_2.log)
tail -5 gpurun_out/r02_pytest_gpu_2.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_2.log || exit 1
SQMC_BUILD_PROFILE=1 timeout 1200 python scripts/bundle_inproc.py 10000000 hci > gpurun_out/r02_bundle_ab.log 2> gpurun_out/r02_bundle_ab.err
echo "ab rc=$?"
cat gpurun_out/r02_bundle_ab.log | cut -c1-200
grep "sqmc build" gpurun_out/r02_bundle_ab.err | tail -40
