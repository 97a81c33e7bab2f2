for n in 131072 1048576; do
  timeout 900 python bench.py --workload sp --seq $n --steps 3 --warmup 3 > gpurun_out/r2_sp1_$n.json 2> gpurun_out/r2_sp1_$n.err || tail -5 gpurun_out/r2_sp1_$n.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_sp1_$n.json').read().strip().splitlines()[-1]); print('$n', round(d['value']/1e6,3), 'M tok/s', round(d['ms_per_step'],2), 'ms', d['kernels'])
except Exception as e:
    print('$n failed', e); print(open('gpurun_out/r2_sp1_$n.err').read()[-1500:])"
done
