set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 tests/run_multi_gpu_parity.py > gpurun_out/r02d_multi_gpu_parity_2gpu.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02d_multi_gpu_parity_2gpu.log)
tail -6 gpurun_out/r02d_multi_gpu_parity_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02d_bench_2gpu.jsonl 2> gpurun_out/r02d_bench_2gpu.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02d_bench_2gpu.jsonl") if l.startswith("{")][-1])
print(round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "build", round(d["build"]["seconds_wall"],3), "parity", d["parity"]["ok"], d["parity"].get("x_dot_y"))
PY
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -2
