for i in 1 2 3; do
for v in v2 v1; do FASTMAX_INTRA=$v python bench.py --no-e2e --no-cpu --steps 10 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernels']
print('$v', round(d['value']), round(d['ms_per_step'],3), 'kernel sum', round(sum(v['ms'] for v in k.values()),3), 'fwd-intra', k.get('mom_seg_fwd',k.get('mom_intra_fwd'))['ms'], 'apply_fwd', k['mom_apply_fwd']['ms'], d['clocks']['sm_mhz'])
"; done; done
