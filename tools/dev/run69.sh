( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 )
for i in 1 2; do python bench.py --no-e2e --no-cpu --steps 10 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernels']
print(round(d['value']), round(d['ms_per_step'],3), 'kernel sum', round(sum(v['ms'] for v in k.values()),3), {n:v['ms'] for n,v in k.items()}, d['clocks']['sm_mhz'])
"; done
