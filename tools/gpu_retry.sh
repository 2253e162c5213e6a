#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <logfile> <command...>: call gpurun until the pod has a free slot
to=$1; log=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 90; else break; fi
done
