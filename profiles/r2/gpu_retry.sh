#!/bin/bash
# gpu_retry.sh <timeout> <out-file> <command...>: gpurun with retries while the pod answers busy (exit code 3)
T=$1; O=$2; shift 2
for try in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $O 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $O; then exit $rc; fi
  sleep 60
done
exit 3
