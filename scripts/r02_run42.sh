set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_host_cpp.py tests/test_gpu_edge_cases.py -x -q > gpurun_out/r02_pytest_gpu_42.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_42.log)
tail -15 gpurun_out/r02_pytest_gpu_42.log
