( timeout 1500 python -m pytest tests/test_gqa_prologue.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 )
