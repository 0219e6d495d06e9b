#!/bin/bash
# gpurun with retries while the pod answers "no box / slot free" (exit code 3, nothing charged).
# usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
