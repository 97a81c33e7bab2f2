P1_P=2 python tools/dev/p1time.py
P1_P=2 python tools/dev/p1time.py
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(seeded and moments and bf16) or golden or headline" 2>&1 | tail -2 )
