set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_incremental.py tests/test_gpu_edge_cases.py tests/test_gpu_configs.py tests/test_gpu_select.py -x -q > gpurun_out/r02_pytest_gpu_35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_35.log)
tail -4 gpurun_out/r02_pytest_gpu_35.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_35.log || exit 1
timeout 600 python scripts/build_ab.py 10000000 hci SQMC_CONNECT_STAGE=1 SQMC_CONNECT_STAGE=1 SQMC_CONNECT_STAGE=1 > gpurun_out/r02_build_ab3.log 2>&1
echo "ab rc=$?"; grep '^{' gpurun_out/r02_build_ab3.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['variant'], round(d['wall_s'], 3), d['phases'], d['same_as_first'])
"
tail -3 gpurun_out/r02_build_ab3.log | cut -c1-300
