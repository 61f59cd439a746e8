#!/bin/bash
# usage: gpurun_retry.sh <tries> <gpurun args...>   -- repeats a gpurun call while the pod answers "busy" (exit 3)
tries=$1; shift
for i in $(seq 1 "$tries"); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i of $tries"; sleep 150
done
exit 3
