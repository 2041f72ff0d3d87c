#!/bin/bash
# usage: [GPURUN_ARGS="--gpus 2"] tools/gpu_retry.sh <logfile> <timeout_s> <command...>
# retries while the pod has no free GPU slot (exit 3)
log=$1; shift; to=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun $GPURUN_ARGS --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 60
done
exit 3
