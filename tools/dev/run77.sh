P1_P=2 python tools/dev/p1time.py
P1_P=2 python tools/dev/p1time.py
P1_P=1 python tools/dev/p1time.py
