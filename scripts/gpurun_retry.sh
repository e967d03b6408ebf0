#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> <logfile> <command...>   -- retries while the pod answers "busy" (exit 3)
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then echo "gpurun rc=$rc (attempt $i)" >> $log; exit $rc; fi
  sleep 120
done
echo "gave up" >> $log
