nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
for e in 0 1 2 4 8 15; do
  L=$PWD/.variants/exp$e/libfastmax_b200.so; 
  FASTMAX_APPLY=gen FASTMAX_LIB=$L timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/tmp/err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exp$e', round(d['ms_per_step'],3), d['clocks'], {k:x['ms'] for k,x in d['kernels'].items() if 'apply' in k})
except Exception as e:
    print('failed', e); print(open('/tmp/err').read()[-600:])"
done
