( timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 )
python tools/dev/p1time.py
