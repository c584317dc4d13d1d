#!/bin/bash
# usage: scripts/gpurun_retry_n.sh <gpus> <timeout-seconds> <command...>  -- retries while the pod answers "busy" (exit 3)
g=$1; t=$2; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
