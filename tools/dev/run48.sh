FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 120 python tools/dev/trace_seg.py > gpurun_out/trace_seg.txt 2>&1
