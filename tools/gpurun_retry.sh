#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun options] -- 'command'     (development aid)
# retries while the pod answers "busy" (exit code 3: nothing charged), every 90 s, for up to ~40 minutes
log=$1; shift
for i in $(seq 1 28); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
