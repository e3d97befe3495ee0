#!/bin/bash
# usage: gpurun_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers "busy" (exit 3)
out=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1; rc=$?
  echo "gpurun exit $rc (attempt $i)" >> "$out"
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
