( FASTMAX_APPLY=gen timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "moments or long_sequence or kept_forward or narrow or headline" 2>&1 | tail -4 )
for e in 0 1 15; do
  L=$PWD/.variants/exp$e/libfastmax_b200.so; [ $e = 0 ] && L=$PWD/fastmax_experiments_b200/libfastmax_b200.so
  FASTMAX_APPLY=gen FASTMAX_LIB=$L timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/tmp/err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exp$e', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:x['ms'] for k,x in d['kernels'].items() if 'apply' in k})
except Exception as e:
    print('failed', e); print(open('/tmp/err').read()[-900:])"
done
