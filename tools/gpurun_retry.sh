#!/bin/bash
# tools/gpurun_retry.sh [gpurun options] -- '<command>'
# gpurun answers rc 3 ("no box or slot free right now", nothing charged) when the pod is busy: wait and ask again.
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt: busy, retrying in 90 s" >&2
  sleep 90
done
exit 3
