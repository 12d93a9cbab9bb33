#!/bin/bash
# usage: gpurun_retry.sh <gpus> <timeout_s> <command...>   -- retries while the pod answers "busy" (rc 3)
G=$1; T=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > /tmp/gpurun_last.txt 2>&1; rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_last.txt; exit $rc
