for i in 1 2 3; do python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], sum(v['ms'] for v in d['kernels'].values()), d['clocks'])
print({k:v['ms'] for k,v in d['kernels'].items()})
"; done
FASTMAX_INTRA=v1 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('v1', d['value'], d['ms_per_step'], sum(v['ms'] for v in d['kernels'].values()), d['clocks'])
"
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
