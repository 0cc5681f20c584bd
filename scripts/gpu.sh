#!/bin/bash
# One parametrised runner for the GPU box: scripts/gpu.sh [--gpus N] [--timeout S] [--tag NAME] -- '<command>'
# Retries while the pod answers "busy" (exit 3, nothing charged); the verdict of the call lands in
# gpurun_out/<tag>.call.txt and whatever the command wrote under gpurun_out/ is merged back.
GPUS=1; TMO=900; TAG=call; TRIES=40
while [ $# -gt 0 ]; do
  case "$1" in
    --gpus) GPUS=$2; shift 2;;
    --timeout) TMO=$2; shift 2;;
    --tag) TAG=$2; shift 2;;
    --tries) TRIES=$2; shift 2;;
    --) shift; break;;
    *) break;;
  esac
done
mkdir -p gpurun_out
ARGS="--timeout $TMO"
[ "$GPUS" != "1" ] && ARGS="$ARGS --gpus $GPUS"
for i in $(seq 1 $TRIES); do
  /usr/local/graft/bin/gpurun $ARGS -- "$1" > gpurun_out/$TAG.call.txt 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc try=$i" >> gpurun_out/$TAG.call.txt; exit $rc; fi
  sleep 90
done
echo "rc=3 gave up after $TRIES tries" >> gpurun_out/$TAG.call.txt
exit 3
