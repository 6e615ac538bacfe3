#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'
# retries while the pod answers "no slot" (exit 3) or "another call of this repo is running" (exit 2 + that message)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  rc=$?
  echo "$out" | tail -60
  if [ $rc -eq 3 ]; then sleep 90; continue; fi
  if [ $rc -eq 2 ] && echo "$out" | grep -q "already running"; then sleep 60; continue; fi
  exit $rc
done
exit 3
