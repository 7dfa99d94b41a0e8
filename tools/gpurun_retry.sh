#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script> [extra gpurun args]; retries while the pod answers busy (rc 3)
T=$1; S=$2; shift 2
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" -- "bash $S"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
