set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_4.log)
tail -15 gpurun_out/r02_pytest_gpu_4.log
grep -q "rc=0" gpurun_out/r02_pytest_gpu_4.log || exit 1
SQMC_BUILD_PROFILE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1e7_run4.json 2> gpurun_out/r02_bench_1e7_run4.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench_1e7_run4.json | cut -c1-3000
grep "sqmc build" gpurun_out/r02_bench_1e7_run4.err | tail -14
