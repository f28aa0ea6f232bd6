#!/bin/bash
# Retries a gpurun call while the pod answers "busy" (exit code 3: nothing charged).  Usage: gpurun_retry.sh <timeout> '<command>'
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
    /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpurun_retry] attempt $attempt: busy, retrying in 60 s"
    sleep 60
done
exit 3
