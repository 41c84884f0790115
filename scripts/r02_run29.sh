set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_bundle.py tests/test_gpu_incremental.py tests/test_gpu_local.py tests/test_gpu_parity.py -x -q > gpurun_out/r02_pytest_gpu_29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_29.log)
tail -4 gpurun_out/r02_pytest_gpu_29.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_29.log || exit 1
timeout 600 python scripts/bundle_inproc.py 10000000 hci 4:44,4:34,4:44,4:34,4:44,4:34 > gpurun_out/r02_bundle_ab12.log 2>&1
grep '^{' gpurun_out/r02_bundle_ab12.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 3), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
timeout 300 python scripts/bundle_inproc.py 200000 lowest 4:44,4:34,4:44,4:34 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 4), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
