set -x
cd $GRAFT_REPO_ROOT
NP=${NP:-2}
run() { # port out extra...
  local port=$1 out=$2; shift 2
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $port bench.py --gpus $NP --steps 20 --warmup 5 "$@" > gpurun_out/$out.jsonl 2> gpurun_out/$out.err
  echo "rc=$? $out"
  python - "$out" <<'PY'
import json,sys
for ln in open("gpurun_out/%s.jsonl"%sys.argv[1]):
    if not ln.startswith("{"): continue
    d=json.loads(ln)
    print(sys.argv[1], d["config"]["workload"][:40], d["config"]["n_dets"], {k:d.get(k) for k in ("n_gpus","ms_per_step")}, "frac",round(d["roofline"]["frac"],3), "e2e_ms", round(d["e2e"]["ms_per_step"],3), d["config"]["vector_exchange"], "build", round(d["build"]["seconds_wall"],3), "parity", d.get("parity",{}).get("ok"), d.get("parity",{}).get("x_dot_y"), "imb", round(d["nnz_per_rank_max_over_mean"],4))
PY
}
if [ "$NP" != "4" ]; then
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29731 tests/run_multi_gpu_parity.py > gpurun_out/r02c_multi_gpu_parity_${NP}gpu.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02c_multi_gpu_parity_${NP}gpu.log)
tail -6 gpurun_out/r02c_multi_gpu_parity_${NP}gpu.log
fi
run 29732 r02c_bench_${NP}gpu
run 29737 r02c_sweep_${NP}gpu --config sweep --geometries 1.0,1.24253,2.0
if [ "$NP" != "8" ]; then
run 29735 r02c_hubbard_${NP}gpu --config hubbard
run 29736 r02c_heg_${NP}gpu --config heg
fi
