for v in ${VARIANTS:-0 1}; do
  SQMC_WCSR=$v python bench.py --n-dets ${NDETS:-10000000} --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('wcsr=$v ms %.3f GB/s %.0f frac %.3f launches %d build %.2fs clocks %s'%(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_launches'], d['build']['seconds_wall'], d['clocks']))"
done
