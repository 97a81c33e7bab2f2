set -x
python tools/prof_run.py 8 32 16384 64 1 --bwd --keep > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|build_kernel|intra_kernel|prep_bwd_kernel|transpose_kernel' -c 12 -o gpurun_out/r2_full python tools/prof_run.py 8 32 16384 64 1 --bwd --keep > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
