( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streaming or first_order" 2>&1 | tail -15 )
