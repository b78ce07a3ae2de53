#!/bin/bash
# usage: tools/gpurun_retry_n.sh <tag> <gpus> <timeout_s> <command...>
tag=$1; shift; n=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$to" -- "$@" > gpurun_out/$tag.out 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/$tag.out || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "rc=$rc" >> gpurun_out/$tag.out
