#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun flags] -- 'command'   (retries while the pod answers "busy", exit code 3)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
