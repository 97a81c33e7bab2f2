timeout 900 python tools/dev/gqa_time.py 2>&1 | tail -70
python bench.py --workload model --model tiny-llama-1.1b --steps 3 --warmup 3 > gpurun_out/r2_model_tiny-llama-1.1b.json 2> gpurun_out/m1.err; tail -c 1500 gpurun_out/r2_model_tiny-llama-1.1b.json
python bench.py --workload model --model tiny-llama-1.1b --steps 3 --warmup 3 --fused > gpurun_out/r2_model_tiny-llama-1.1b_fused.json 2> gpurun_out/m2.err; tail -c 1500 gpurun_out/r2_model_tiny-llama-1.1b_fused.json; tail -3 gpurun_out/m2.err
