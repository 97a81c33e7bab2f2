( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "p1 or (seeded and moments) or linearmax or autograd or variants or rect or kept" 2>&1 | tail -15 )
P1_CAUSAL=0 python tools/dev/p1time.py
