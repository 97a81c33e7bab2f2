for n in 1 2 3 4 5; do
  FASTMAX_LIB=$PWD/.variants/cfg$n/libfastmax_b200.so timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/tmp/err_$n | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', round(d['ms_per_step'],3), {k:x['ms'] for k,x in d['kernels'].items() if 'apply' in k or 'build' in k})
except Exception as e:
    print('$n failed', e); print(open('/tmp/err_$n').read()[-800:])"
done
