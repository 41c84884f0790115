set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/bundle_inproc.py 200000 lowest 4:34,4:134,4:141,4:142,4:144,4:148 > gpurun_out/r02_bundle_ab9_small.log 2>&1
echo "small rc=$?"; grep '^{' gpurun_out/r02_bundle_ab9_small.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 4), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
grep -q '"kernel": 148' gpurun_out/r02_bundle_ab9_small.log || { tail -5 gpurun_out/r02_bundle_ab9_small.log; exit 1; }
timeout 600 python scripts/bundle_inproc.py 10000000 hci 4:34,4:134,4:141,4:142,4:144,4:148,4:34,4:142,4:144 > gpurun_out/r02_bundle_ab9.log 2>&1
echo "ab rc=$?"; grep '^{' gpurun_out/r02_bundle_ab9.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 3), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
