timeout 600 python tools/dev/gqa_dbg.py 2>&1 | tail -40
