mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_model_level.py -x -q -m gpu -k "d128 or 128 or head_size or narrow or golden or model" 2>&1 | tail -6 )
python bench.py --workload model --model Llama-2-7b-hf > gpurun_out/r2_model_llama.json 2> gpurun_out/model.err; tail -2 gpurun_out/model.err; cut -c1-1200 gpurun_out/r2_model_llama.json
