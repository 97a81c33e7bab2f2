D128_BWD=1 timeout 600 python tools/dev/d128.py 2>&1 | tail -12
