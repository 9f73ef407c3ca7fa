#!/bin/bash
# gpurun with retries on "no slot" (exit code 3): tools/gpurun_retry.sh <log> <timeout> <command...>
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
