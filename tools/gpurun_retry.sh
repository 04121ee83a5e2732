#!/bin/bash
# gpurun with retries while the pod answers busy/transient (nothing is charged for those).  usage: gpurun_retry.sh LOG [gpurun args...]
LOG=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then exit 0; fi
  sleep 120
done
