P1_P=2 python tools/dev/p1time.py
P1_P=2 FASTMAX_LIB=$PWD/.variants/flags/libfastmax_b200.so python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/flags/libfastmax_b200.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "headline or (seeded and moments and 2560)" 2>&1 | tail -3
