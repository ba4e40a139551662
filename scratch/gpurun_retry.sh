#!/bin/bash
# usage: scratch/gpurun_retry.sh LOG TIMEOUT 'command' [gpus]   -- retries while the pod answers busy (rc 3)
log=$1; to=$2; cmd=$3; gp=${4:-1}
for i in $(seq 1 20); do
  if [ "$gp" = 1 ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$cmd" > $log 2>&1; rc=$?
  else /usr/local/graft/bin/gpurun --gpus $gp --timeout $to -- "$cmd" > $log 2>&1; rc=$?; fi
  if [ $rc != 3 ]; then break; fi
  sleep 150
done
echo "done rc=$rc" >> $log
