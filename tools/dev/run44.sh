( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(seeded and moments and bf16) or headline or long_sequence or kept or graph or golden or narrow" 2>&1 | tail -6 )
P1_P=2 python tools/dev/p1time.py
