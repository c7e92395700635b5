#!/bin/bash
# usage: scratch/gpu_retry.sh <tag> <timeout_s> [--gpus N] -- <command>
# retries a gpurun call while the pod answers "busy/transient" (exit 3), nothing is charged for those
tag=$1; shift; to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > gpurun_out/$tag.out 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/$tag.out; then break; fi
  sleep 60
done
echo "rc=$rc attempts=$i"
tail -60 gpurun_out/$tag.out
