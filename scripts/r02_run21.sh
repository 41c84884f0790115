set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_pt2.py -x -q > gpurun_out/r02_pytest_gpu_21a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_21a.log)
tail -30 gpurun_out/r02_pytest_gpu_21a.log
(timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_pt2.py > gpurun_out/r02_pytest_gpu_21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_21.log)
tail -6 gpurun_out/r02_pytest_gpu_21.log
