#!/bin/bash
# retry a gpurun command while the pod answers "transient" (nothing charged); usage: gpurun_retry.sh <tries> <cmd>
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(timeout 3000 /usr/local/graft/bin/gpurun --timeout 1500 -- "$@" 2>&1)
  echo "$out" | tail -25
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  sleep 90
done
exit 3
