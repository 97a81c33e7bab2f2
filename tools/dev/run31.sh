mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 )
python tools/sweep.py > gpurun_out/r2_sweep_table.md 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err
python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; wc -c gpurun_out/r2_bench_full.json
