#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient: nothing is charged)
# usage: scripts/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then
    sleep 45
    continue
  fi
  echo "$out"
  exit $rc
done
echo "gpurun_retry: still busy after 40 attempts"
exit 3
