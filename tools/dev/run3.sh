for e in 0 1 2 4 3; do echo "=== FM_EXP=$e"; FASTMAX_LIB=$PWD/.variants/exp$e/libfastmax_b200.so timeout 120 python tools/timing.py 2>&1 | grep -v "^ *{}" | tail -4; done
echo "=== bwd exp0"; FASTMAX_LIB=$PWD/.variants/exp0/libfastmax_b200.so timeout 120 python tools/timing.py --bwd 2>&1 | grep "apply\|build"
echo "=== full build exp0"; FASTMAX_BUILD=full FASTMAX_LIB=$PWD/.variants/exp0/libfastmax_b200.so timeout 120 python tools/timing.py 2>&1 | grep "build"
for b in packed full; do
  FASTMAX_BUILD=$b FASTMAX_APPLY=old timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$b', round(d['ms_per_step'],3), {k:(x['launches'], x['ms']) for k,x in d['kernels'].items() if 'build' in k})"
done
