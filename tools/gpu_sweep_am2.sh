#!/bin/bash
# AM/SSB: FIR worker warps per SM beyond 48 (finer shares: smaller tail, more warm-up tiles)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sweep_am2.txt
for wl in am ssb; do
  for w in 40 48 60 72 80 96 120; do
    SDR_AM_WARPS_PER_SM=$w timeout 200 python bench.py --workload $wl --steps 3000 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$wl warps/SM $w:', d['value'], 'Msps  frac', d['roofline']['frac'], ' ms', d['ms_per_step'])
except Exception as ex:
    print('$wl $w bench failed', ex)
" | tee -a gpurun_out/sweep_am2.txt
  done
done
