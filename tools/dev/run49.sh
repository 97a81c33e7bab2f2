( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(seeded and moments and bf16) or golden" 2>&1 | tail -3 )
P1_P=2 python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/trace/libfastmax_b200.so timeout 120 python tools/dev/trace_seg.py > gpurun_out/trace_seg.txt 2>&1
