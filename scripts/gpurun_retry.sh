#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <logfile> [--gpus N] -- <command>: retries while the pod answers "busy" (exit code 3)
t=$1; log=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $t "$@" > $log 2>&1; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
