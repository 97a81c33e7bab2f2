# same-box A/B: the in-tree library against .variants/head (built from the last commit)
for i in 1 2; do
P1_P=2 python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/head/libfastmax_b200.so P1_P=2 python tools/dev/p1time.py
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(seeded and moments and bf16) or golden or headline" 2>&1 | tail -2
