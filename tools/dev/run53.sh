( timeout 1500 python -m pytest tests/test_gqa_prologue.py -x -q -m gpu 2>&1 | tail -25 )
