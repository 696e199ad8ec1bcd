#!/bin/bash
# NBFM kernel: worker warps per SM (SDR_FM_WARPS_PER_SM) sweep, tone (tensor-core tuner) and noise (SIMT tuner)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sweep_fm.txt
for sig in tone noise; do
  for w in 16 32 48 96; do
    SDR_FM_WARPS_PER_SM=$w timeout 200 python bench.py --workload fm --signal $sig --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('fm $sig warps/SM $w:', d['value'], 'Msps  frac', d['roofline']['frac'], ' ms', d['ms_per_step'])
except Exception as ex:
    print('fm $sig $w bench failed', ex)
" | tee -a gpurun_out/sweep_fm.txt
  done
done
