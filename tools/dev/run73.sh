FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 300 python tools/dev/trace_apply2.py 2>&1 | tail -36
