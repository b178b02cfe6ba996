#!/bin/bash
# usage: scripts/gpurun_retry.sh LOGFILE TIMEOUT 'command'  -- gpurun with retries while the pod answers "transient" (exit code 3)
LOG=$1; TO=$2; CMD=$3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$CMD" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 45
done
exit 3
