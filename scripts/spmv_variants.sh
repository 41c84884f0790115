for v in ${VARIANTS:-0 6 7 8 9}; do
  SQMC_SPMV_VARIANT=$v python bench.py --n-dets ${NDETS:-10000000} --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('variant $v ms %.3f GB/s %.0f frac %.3f clocks %s'%(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks']))"
done
