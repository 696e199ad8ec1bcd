#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "== $cfg" | tee -a gpurun_out/probe_timeline.txt
  env $cfg timeout 200 python tools/probe_timeline.py 2>&1 | tee -a gpurun_out/probe_timeline.txt
done
