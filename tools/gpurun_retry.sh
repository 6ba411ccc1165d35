#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   — retries while the pod answers "transient" / busy (rc 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1; rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
tail -100 $log
