N=$1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -3 gpurun_out/r2_bench_${N}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('weak', d['value']/1e6, d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'))
sp=d['sp']; print('sp', sp['value']/1e6, sp['ms_per_step'], sp['parity'], sp['efficiency'], sp['exchange'])"
