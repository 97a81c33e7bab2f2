#!/bin/bash
# usage: tools/dev/gpu_retry.sh <out-file> <timeout-s> <command...>   (retries while the pod answers "transient")
out=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $out 2>&1
  if ! grep -q "status=transient" $out; then exit 0; fi
  sleep 90
done
