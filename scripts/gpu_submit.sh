#!/bin/bash
# Submits a command to gpurun and resubmits while the pod answers "busy" (exit 3).  Usage: scripts/gpu_submit.sh <timeout_s> '<command>'
to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
