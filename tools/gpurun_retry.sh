#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod is busy
log="$1"; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
exit 3
