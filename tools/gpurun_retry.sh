#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit 3 or status=transient: nothing charged).
# Usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|no box or slot"; then sleep 60; continue; fi
  if [ $rc -eq 3 ]; then sleep 60; continue; fi
  echo "$out" | tail -n 120; exit $rc
done
echo "gave up after 40 attempts"; exit 3
