#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'  [gpus]
# retries while the pod answers "busy" (exit 3); logs to gpurun_out/_retry.log
T=$1; CMD=$2; G=${3:-1}
mkdir -p /root/repo/gpurun_out
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD" > /root/repo/gpurun_out/_retry.log 2>&1; rc=$?
  else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" > /root/repo/gpurun_out/_retry.log 2>&1; rc=$?; fi
  if [ $rc -ne 3 ]; then echo "rc=$rc after $i tries"; tail -60 /root/repo/gpurun_out/_retry.log; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
