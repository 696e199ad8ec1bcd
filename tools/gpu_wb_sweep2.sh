#!/bin/bash
# WBFM gen-2 sweep: which warp runs the recurrence (SDR_WB_REC), channels per CTA (SDR_WB_G)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --workload wbfm --no-extras --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$cfg wbfm', d['value'], 'Msps  frac', d['roofline']['frac'], ' ms', d['ms_per_step'])
except Exception as ex:
    print('$cfg bench failed', ex)
" | tee -a gpurun_out/wb_sweep2.txt
done
