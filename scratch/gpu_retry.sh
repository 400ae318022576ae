#!/bin/bash
# usage: scratch/gpu_retry.sh <log> <timeout> <command...>   -- retries a gpurun call while the pod answers "transient" / busy (nothing charged)
LOG=$1; TMO=$2; shift 2
for try in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -5 $LOG
