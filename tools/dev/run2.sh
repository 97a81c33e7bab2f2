set -x
# 1. parity of the packed build (default) + gen4 apply, a focused subset
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2b_tests.log 2>&1
# 2. step breakdown: packed vs full build
for b in packed full; do
  FASTMAX_BUILD=$b FASTMAX_APPLY=old timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_bench_$b.json 2> gpurun_out/r2b_bench_$b.err
done
# 3. in-kernel counters of the generated-operand apply
FASTMAX_LIB=$PWD/.variants/timing/libfastmax_b200.so timeout 300 python tools/timing.py --bwd > gpurun_out/r2b_timing.log 2>&1
cat gpurun_out/r2b_tests.log
for v in packed full; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), {k:(x["launches"], x["ms"]) for k,x in d["kernels"].items()})
except Exception as e:
    print("$v", "failed", e); print(open("gpurun_out/r2b_bench_$v.err").read()[-1500:])
PY
done
cat gpurun_out/r2b_timing.log
