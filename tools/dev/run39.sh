timeout 900 python tools/dev/split.py 2>&1 | tail -14
