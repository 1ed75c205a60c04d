#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <logfile> <command...>   -- retries while the pod answers "busy" (exit 3)
t=$1; log=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
