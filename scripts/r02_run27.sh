set -x
cd $GRAFT_REPO_ROOT
timeout 600 python scripts/bundle_inproc.py 10000000 hci 4:141,4:151,4:152,4:153,4:141,4:151,4:152 > gpurun_out/r02_bundle_ab10.log 2>&1
echo "ab rc=$?"; grep '^{' gpurun_out/r02_bundle_ab10.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 3), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
tail -2 gpurun_out/r02_bundle_ab10.log | cut -c1-200
