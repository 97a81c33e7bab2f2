mkdir -p gpurun_out
P1_CAUSAL=0 ncu --set full --clock-control none --import-source on -k regex:"reduce_kernel|applyfull_kernel|finalize_kernel|prefix_kernel" -c 9 -s 18 -o gpurun_out/r2_p1full -f python tools/dev/p1time.py > gpurun_out/ncu_p1full.log 2>&1; tail -2 gpurun_out/ncu_p1full.log
