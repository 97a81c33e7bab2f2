mkdir -p gpurun_out
python tools/sweep.py > gpurun_out/r2_sweep_table.md 2> gpurun_out/sweep.err; tail -3 gpurun_out/sweep.err
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -c 4 -s 8 -o gpurun_out/r2_p1scan -f python tools/dev/p1time.py > gpurun_out/ncu_p1.log 2>&1; tail -3 gpurun_out/ncu_p1.log
