set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/bundle_inproc.py 200000 lowest 4:34,4:624,4:634,4:625,4:626,4:534,4:544 > gpurun_out/r02_bundle_ab5_small.log 2>&1
echo "small rc=$?"; grep '^{' gpurun_out/r02_bundle_ab5_small.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 4), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
tail -3 gpurun_out/r02_bundle_ab5_small.log | cut -c1-300
grep -q '"kernel": 544' gpurun_out/r02_bundle_ab5_small.log || exit 1
timeout 600 python scripts/bundle_inproc.py 10000000 hci 4:34,4:624,4:634,4:644,4:625,4:635,4:626,4:636,4:524,4:534,4:544,4:34 > gpurun_out/r02_bundle_ab5.log 2>&1
echo "ab rc=$?"; grep '^{' gpurun_out/r02_bundle_ab5.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['R'], d['kernel'], round(d['ms'], 3), round(d['frac_6455.6'], 4), d['max_rel_diff_vs_first'])
"
tail -3 gpurun_out/r02_bundle_ab5.log | cut -c1-300
