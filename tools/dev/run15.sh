( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "recurrent" 2>&1 | tail -8 )
