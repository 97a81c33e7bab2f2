timeout 900 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; tail -2 gpurun_out/r2_bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 900 python tools/sweep.py > gpurun_out/r2_sweep.md 2> gpurun_out/r2_sweep.err; tail -2 gpurun_out/r2_sweep.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_full.json').read().strip().splitlines()[-1])
print(d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['dominant_kernel'], d['e2e'], d['cpu_baseline'], d['gpu_launches'])
r=json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1]); print(r['value'], r['cpu_baseline'])
"
cat gpurun_out/r2_sweep.md
