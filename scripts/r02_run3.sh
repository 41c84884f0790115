set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
export NCCL_DEBUG=WARN
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/run_multi_gpu_parity.py > gpurun_out/r02_multi_gpu_parity_2gpu.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02_multi_gpu_parity_2gpu.log)
tail -12 gpurun_out/r02_multi_gpu_parity_2gpu.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --n-dets 2000000 --steps 10 --warmup 3 > gpurun_out/r02_bench_2e6_2gpu.json 2> gpurun_out/r02_bench_2e6_2gpu.err; echo "bench rc=$?")
tail -c 2500 gpurun_out/r02_bench_2e6_2gpu.json | cut -c1-2500
tail -5 gpurun_out/r02_bench_2e6_2gpu.err
(SQMC_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --n-dets 2000000 --steps 10 --warmup 3 --no-parity > gpurun_out/r02_bench_2e6_2gpu_nccl.json 2> gpurun_out/r02_bench_2e6_2gpu_nccl.err; echo "bench nccl rc=$?")
head -c 1200 gpurun_out/r02_bench_2e6_2gpu_nccl.json
