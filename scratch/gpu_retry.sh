#!/bin/bash
# usage: scratch/gpu_retry.sh <log> <timeout> [--gpus N] <command>   -- retries a gpurun call while the pod answers "transient" / busy (nothing charged)
LOG=$1; TMO=$2; shift 2
OPTS=""
if [ "$1" == "--gpus" ]; then OPTS="--gpus $2"; shift 2; fi
for try in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TMO $OPTS -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -5 $LOG
