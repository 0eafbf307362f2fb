#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pod answers "transient" / busy, nothing charged)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|no box or slot\|retry in a few minutes"; then
    sleep 45
    continue
  fi
  echo "$out"
  exit 0
done
echo "$out"
echo "gpurun_retry: gave up"
exit 3
