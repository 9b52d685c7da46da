#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 = nothing charged).  usage: scripts/grun.sh <timeout> [--gpus N] -- '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  gpurun --timeout $T "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
