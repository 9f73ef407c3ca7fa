#!/bin/bash
# gpurun with retries on "no slot" (exit code 3): [GPUS=N] tools/gpurun_retry.sh <log> <timeout> <command...>
log=$1; to=$2; shift 2
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 80); do
  /usr/local/graft/bin/gpurun $extra --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 40
done
exit 3
