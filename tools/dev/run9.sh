( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 )
FASTMAX_LIB=$PWD/.variants/timing/libfastmax_b200.so timeout 120 python tools/timing.py 2>&1 | grep "build\|special\|whole"
FASTMAX_BUILD=full FASTMAX_LIB=$PWD/.variants/timing/libfastmax_b200.so timeout 120 python tools/timing.py 2>&1 | grep "build\|special\|whole"
