( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gqa_prologue.py -x -q -m gpu -k "(seeded and moments and bf16) or golden or headline or long_sequence or kept or gqa_matches or narrow" 2>&1 | tail -4 )
P1_P=2 python tools/dev/p1time.py
