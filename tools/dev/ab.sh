# same-box A/B: the in-tree library against .variants/head
for i in 1 2; do
P1_P=2 python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/head/libfastmax_b200.so P1_P=2 python tools/dev/p1time.py
done
python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default bench', d['value'], d['ms_per_step'], d['clocks'])"
FASTMAX_LIB=$PWD/.variants/head/libfastmax_b200.so python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('head bench', d['value'], d['ms_per_step'], d['clocks'])"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(seeded and moments and bf16) or golden or headline" 2>&1 | tail -2
