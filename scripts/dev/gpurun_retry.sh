#!/bin/sh
# usage: gpurun_retry.sh <timeout_s> <gpus> '<command>'   -- retries while the pod answers "busy" (exit 3)
t=$1; g=$2; shift 2
i=0
while [ $i -lt 40 ]; do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  i=$((i+1)); echo "[retry $i] busy, sleeping 90 s"; sleep 90
done
exit 3
