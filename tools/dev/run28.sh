P1_CAUSAL=0 python tools/dev/p1time.py
