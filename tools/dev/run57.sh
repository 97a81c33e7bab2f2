( timeout 900 python -m pytest tests/test_gqa_prologue.py -x -q -m gpu 2>&1 | tail -4 )
timeout 900 python tools/dev/gqa_time.py 2>&1 | tail -18
