#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy / transient" (exit 3, nothing charged).
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout-seconds> '<command>'
T=$1; shift
G=${GPUS:-1}
for attempt in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 90 s" >&2
  sleep 90
done
exit 3
