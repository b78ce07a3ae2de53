#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> <timeout_s> <command...>   (retries while the pod answers "busy / transient")
tag=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > gpurun_out/$tag.out 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/$tag.out || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "rc=$rc" >> gpurun_out/$tag.out
