FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 120 python tools/dev/trace_seg.py 2>&1 | tail -34
