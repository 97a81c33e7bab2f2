mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; wc -c gpurun_out/r2_bench_full.json
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
python tools/prof_run.py 8 32 16384 64 1 --bwd --keep > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|build_kernel|intra_kernel|fwd_kernel|dkdv_kernel|dq_kernel|prep_bwd_kernel|transpose_kernel' -c 14 -o gpurun_out/r2_full -f python tools/prof_run.py 8 32 16384 64 1 --bwd --keep > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
