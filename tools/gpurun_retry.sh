#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- <command>   (retries while the pod answers busy/transient)
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t "$@" 2>&1)
  if echo "$out" | grep -qE "status=transient|rc=3|exit code 3|answers busy|no box"; then
    sleep 120
    continue
  fi
  echo "$out"
  exit 0
done
echo "gave up: $out"
