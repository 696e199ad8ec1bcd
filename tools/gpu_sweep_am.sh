#!/bin/bash
# AM/SSB: FIR worker warps per SM (SDR_AM_WARPS_PER_SM) x recurrence helper warps (SDR_DC_HELPERS)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sweep_am.txt
for wl in am ssb mixed; do
  for h in 3 7; do
    for w in 16 24 32 48; do
      SDR_DC_HELPERS=$h SDR_AM_WARPS_PER_SM=$w timeout 200 python bench.py --workload $wl --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$wl helpers $h warps/SM $w:', d['value'], 'Msps  frac', d['roofline']['frac'], ' ms', d['ms_per_step'])
except Exception as ex:
    print('$wl $h $w bench failed', ex)
" | tee -a gpurun_out/sweep_am.txt
    done
  done
done
