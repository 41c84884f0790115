set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/bundle_inproc.py 200000 lowest 4:44,4:54,4:44,4:54 > gpurun_out/r02_bundle_ab13_small.log 2>&1
echo "small rc=$?"; grep '^{' gpurun_out/r02_bundle_ab13_small.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 4), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
grep -c '"kernel": 54' gpurun_out/r02_bundle_ab13_small.log || { tail -5 gpurun_out/r02_bundle_ab13_small.log; exit 1; }
timeout 600 python scripts/bundle_inproc.py 10000000 hci 4:44,4:54,4:44,4:54,4:44,4:54 > gpurun_out/r02_bundle_ab13.log 2>&1
echo "ab rc=$?"; grep '^{' gpurun_out/r02_bundle_ab13.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 3), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
