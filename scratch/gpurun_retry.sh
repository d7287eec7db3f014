#!/bin/bash
# scratch/gpurun_retry.sh LOG TIMEOUT 'command'  — gpurun, retried while the pod answers "transient" (nothing charged)
LOG=$1; T=$2; CMD=$3
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  gpurun ${GPUS:+--gpus $GPUS} --timeout $T -- "$CMD" > $LOG 2>&1
  if grep -q "status=transient" $LOG || grep -q "exit code 3" $LOG; then sleep 90; continue; fi
  break
done
