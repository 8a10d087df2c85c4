#!/bin/bash
# usage: tools/gpurun_retry.sh TIMEOUT_S 'command'   -- retries while the pod answers "busy" (exit 3), up to ~60 min
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $i answered busy; sleeping 120 s"; sleep 120
done
exit 3
