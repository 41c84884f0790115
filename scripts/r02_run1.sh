set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_1.log)
tail -5 gpurun_out/r02_pytest_gpu_1.log
(timeout 600 python bench.py --n-dets 1000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1e6_1gpu.json 2> gpurun_out/r02_bench_1e6_1gpu.err; echo "bench rc=$?")
tail -c 1500 gpurun_out/r02_bench_1e6_1gpu.json
tail -5 gpurun_out/r02_bench_1e6_1gpu.err
