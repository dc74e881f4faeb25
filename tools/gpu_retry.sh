#!/bin/bash
# usage: tools/gpu_retry.sh <logname> <gpurun args...>   — retries while the pod answers busy (exit 3), up to 40 times
name=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > gpurun_out/$name.out 2>&1
  rc=$?
  echo "attempt $i rc=$rc" >> gpurun_out/$name.attempts
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/$name.out; then break; fi
  sleep 90
done
echo "done rc=$rc" >> gpurun_out/$name.attempts
