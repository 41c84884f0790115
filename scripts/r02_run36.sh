set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_36.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_36.log)
tail -4 gpurun_out/r02_pytest_gpu_36.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_36.log || exit 1
python -c "import __graft_entry__ as g; g.smoke()"
SQMC_DAV_PROFILE=1 timeout 600 python scripts/davidson_states.py 10000000 > gpurun_out/r02c_davidson_states_1e7.jsonl 2> gpurun_out/r02c_davidson_states_1e7.err
echo "dav rc=$?"; cut -c1-300 gpurun_out/r02c_davidson_states_1e7.jsonl; grep -i "h.v\|matvec" gpurun_out/r02c_davidson_states_1e7.err | tail -6
