mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 )
for m in Llama-2-7b-hf tiny-llama-1.1b pythia-14m; do python bench.py --workload model --model $m > gpurun_out/r2_model_$m.json 2> gpurun_out/model_$m.err; tail -1 gpurun_out/model_$m.err | cut -c1-200; done
python tools/sweep.py > gpurun_out/r2_sweep_table.md 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err
