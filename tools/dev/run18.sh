( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 )
for i in 1 2; do timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/tmp/err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:x['ms'] for k,x in d['kernels'].items()})
except Exception as e:
    print('failed', e); print(open('/tmp/err').read()[-900:])"; done
FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 120 python tools/dev/trace_intra.py 2>&1 | tail -40
