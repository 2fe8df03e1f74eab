#!/bin/bash
# usage: gpurun_retry.sh <timeout> [--gpus N] -- <command>   (retries while the pod answers busy: exit code 3)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
