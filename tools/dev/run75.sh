P1_P=2 python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/iss4/libfastmax_b200.so P1_P=2 python tools/dev/p1time.py
P1_P=2 python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/iss4/libfastmax_b200.so P1_P=2 python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/iss4/libfastmax_b200.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "seeded and moments and bf16" 2>&1 | tail -2
