#!/bin/bash
# gpurun_retry.sh LOG TIMEOUT CMD : retry a gpurun call while the pod answers "transient" (nothing charged)
log=$1; to=$2; shift 2
for i in $(seq 1 20); do
  gpurun --timeout $to ${GPUS:+--gpus $GPUS} -- "$@" > $log 2>&1
  grep -q "status=transient" $log || break
  sleep 90
done
