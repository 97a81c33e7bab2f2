mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/b8.err; tail -3 gpurun_out/b8.err; wc -c gpurun_out/r2_bench_8gpu.json
