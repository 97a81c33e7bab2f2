for m in pythia-14m tiny-llama-1.1b Llama-2-7b-hf; do
  timeout 900 python bench.py --workload model --model $m --steps 3 --warmup 3 > gpurun_out/r2_model_$m.json 2> gpurun_out/r2_model_$m.err || tail -5 gpurun_out/r2_model_$m.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_model_$m.json').read().strip().splitlines()[-1]); print('$m', round(d['value'],1), 'tok/s', round(d['ms_per_step'],2), 'ms', d['attention']['share_of_step'], d['loss'])
except Exception as e:
    print('$m failed', e); print(open('gpurun_out/r2_model_$m.err').read()[-1500:])"
done
