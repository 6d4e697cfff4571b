#!/bin/bash
# usage: gpu_retry.sh <timeout> '<command>'  -- retries gpurun while the pod answers "busy" (exit 3), up to ~40 min
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
