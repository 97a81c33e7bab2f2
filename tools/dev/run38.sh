python tools/dev/sanit.py 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck python tools/dev/sanit.py > gpurun_out/sanitizer.log 2>&1; tail -6 gpurun_out/sanitizer.log
