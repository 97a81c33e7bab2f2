FASTMAX_APPLY=gen FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 120 python tools/dev/trace_apply.py fwd 2>&1 | tail -36
FASTMAX_APPLY=gen FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 120 python tools/dev/trace_apply.py bwd 2>&1 | tail -36
