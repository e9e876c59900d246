#!/bin/bash
# usage: scratch/gpu.sh <tag> <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (exit 3), logs to gpurun_out/<tag>.out
tag=$1; to=$2; shift 2
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > gpurun_out/$tag.out 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> gpurun_out/$tag.out; exit $rc; fi
  sleep 90
done
