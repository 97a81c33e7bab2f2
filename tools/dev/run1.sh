set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( FASTMAX_APPLY=gen4 timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2a_tests_gen4.log 2>&1
for v in gen4 gen2 old; do
  FASTMAX_APPLY=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2a_bench_$v.json 2> gpurun_out/r2a_bench_$v.err
done
( FASTMAX_APPLY=gen2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "moments or long_sequence or kept_forward or narrow or headline" 2>&1 | tail -8 ) > gpurun_out/r2a_tests_gen2.log 2>&1
tail -5 gpurun_out/r2a_tests_gen4.log gpurun_out/r2a_tests_gen2.log
for v in gen4 gen2 old; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2a_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), {k:(x["launches"], x["ms"]) for k,x in d["kernels"].items()})
except Exception as e:
    print("$v", "failed", e); print(open("gpurun_out/r2a_bench_$v.err").read()[-1500:])
PY
done
