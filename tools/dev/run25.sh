python tools/dev/p1time.py
for v in "$@"; do FASTMAX_LIB=$PWD/.variants/$v/libfastmax_b200.so python tools/dev/p1time.py; done
