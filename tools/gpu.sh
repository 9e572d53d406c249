#!/bin/bash
# dev tool: gpurun with retries on "no box free" (rc 3).  usage: [GPUS=2] tools/gpu.sh <timeout_s> '<command>'
t=$1; shift
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$t" $extra -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
