#!/bin/bash
# usage: tools/gpu_retry.sh <gpurun args...>   -- retries while the pod answers "busy / transient" (nothing charged)
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient"; then
    sleep 90
    continue
  fi
  echo "$out"
  exit $rc
done
echo "gpu_retry: gave up after 40 attempts"
exit 3
