#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3); usage: tools/gpu/retry.sh <timeout_s> <command...>
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
