#!/bin/bash
# usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pod answers "busy", exit code 3)
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 90 s" >&2
  sleep 90
done
exit 3
