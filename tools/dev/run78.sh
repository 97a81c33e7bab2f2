for i in 1 2; do
P1_P=2 python tools/dev/p1time.py
FASTMAX_LIB=$PWD/.variants/head/libfastmax_b200.so P1_P=2 python tools/dev/p1time.py
done
