#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3: nothing charged).  usage: gpurun_retry.sh <timeout> <command...>
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 75
done
exit 3
