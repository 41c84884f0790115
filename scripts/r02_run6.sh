set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
nvidia-smi topo -m | head -12
run() { # nproc port out extra...
  local np=$1 port=$2 out=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 20 --warmup 5 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  echo "rc=$? $out"
  python - "$out" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ("n_gpus","ms_per_step","value")}, "frac",d["roofline"]["frac"], "e2e_ms", d["e2e"]["ms_per_step"], d["e2e"]["call"], d["config"]["vector_exchange"], "build", d["build"]["seconds_wall"], "parity", d.get("parity",{}).get("ok"), d.get("parity",{}).get("x_dot_y"), d.get("parity",{}).get("y_norm2"), "imb", d["nnz_per_rank_max_over_mean"])
PY
}
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tests/run_multi_gpu_parity.py > gpurun_out/r02_multi_gpu_parity_8gpu.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02_multi_gpu_parity_8gpu.log)
tail -6 gpurun_out/r02_multi_gpu_parity_8gpu.log
run 8 29622 r02_bench_8gpu
SQMC_P2P=0 run 8 29623 r02_bench_8gpu_nccl --no-parity
run 4 29624 r02_bench_4gpu
run 2 29625 r02_bench_2gpu
