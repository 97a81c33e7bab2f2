mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python tools/sweep.py > gpurun_out/r2_sweep_table.md 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err
python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; wc -c gpurun_out/r2_bench_full.json
