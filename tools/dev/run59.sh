FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 300 python tools/dev/trace_intra.py > gpurun_out/trace_intra.txt 2>&1; tail -3 gpurun_out/trace_intra.txt
