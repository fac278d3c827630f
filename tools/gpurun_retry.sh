#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args...] -- 'command'    retries while the pod answers "no box free" (exit code 3)
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt: no box, sleeping 90 s"
  sleep 90
done
exit 3
