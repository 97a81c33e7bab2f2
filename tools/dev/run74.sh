( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gqa_prologue.py -x -q -m gpu -k "(seeded and moments and bf16) or golden or headline or long_sequence or kept or gqa_matches or narrow" 2>&1 | tail -4 )
P1_P=2 python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 300 python tools/dev/trace_apply2.py 2>&1 | tail -12
