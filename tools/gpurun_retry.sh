#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy/transient
log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" $log && ! grep -q "status=ok\|status=fail" $log; then sleep 150; continue; fi
  break
done
