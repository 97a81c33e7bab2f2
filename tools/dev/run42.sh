D128_BWD=1 timeout 600 python tools/dev/d128.py 2>&1 | tail -7
timeout 900 python tools/dev/split.py 2>&1 | tail -10
FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 120 python tools/dev/trace_d128.py 2>&1 | tail -8
