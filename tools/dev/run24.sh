FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 120 python tools/dev/trace_p1.py 2>&1 | head -40
