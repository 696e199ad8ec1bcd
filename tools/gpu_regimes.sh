#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "$cfg: $(env $cfg timeout 300 python tools/probe_regimes.py 2 2>&1 | tail -1)" | tee -a gpurun_out/probe_regimes.txt
done
