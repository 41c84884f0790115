set -x
cd $GRAFT_REPO_ROOT
run() { # nproc port out extra...
  local np=$1 port=$2 out=$3; shift 3
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 20 --warmup 5 "$@" > gpurun_out/$out.jsonl 2> gpurun_out/$out.err
  echo "rc=$? $out"
  python - "$out" <<'PY'
import json,sys
for ln in open("gpurun_out/%s.jsonl"%sys.argv[1]):
    if not ln.startswith("{"): continue
    d=json.loads(ln)
    print(sys.argv[1], d["config"]["workload"][:40], d["config"]["n_dets"], {k:d.get(k) for k in ("n_gpus","ms_per_step")}, "frac",round(d["roofline"]["frac"],3), "e2e_ms", round(d["e2e"]["ms_per_step"],3), d["config"]["vector_exchange"], "build", round(d["build"]["seconds_wall"],3), "parity", d.get("parity",{}).get("ok"), d.get("parity",{}).get("x_dot_y"), "imb", round(d["nnz_per_rank_max_over_mean"],4))
PY
}
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tests/run_multi_gpu_parity.py > gpurun_out/r02_multi_gpu_parity_8gpu_final.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02_multi_gpu_parity_8gpu_final.log)
tail -5 gpurun_out/r02_multi_gpu_parity_8gpu_final.log
run 8 29632 r02_final_bench_8gpu
run 4 29633 r02_final_bench_4gpu
run 2 29634 r02_final_bench_2gpu
run 8 29635 r02_final_hubbard_8gpu --config hubbard
run 8 29636 r02_final_heg_8gpu --config heg
run 8 29637 r02_final_sweep_8gpu --config sweep --geometries 1.0,1.24253,2.0
