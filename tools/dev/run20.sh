set -x
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
python tools/prof_run.py 8 32 16384 64 1 --bwd --keep > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fm -c 16 -o gpurun_out/r2_full python tools/prof_run.py 8 32 16384 64 1 --bwd --keep > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
