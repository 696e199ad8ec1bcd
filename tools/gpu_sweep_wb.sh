#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "14 2" "19 2" "19 0" "19 4" "21 2" "21 1"; do
  set -- $cfg
  SDR_WB_G=$1 SDR_WB_S3=$2 timeout 300 python bench.py --workload wbfm --no-extras --no-cpu --steps 40 2>&1 | tail -1 | python -c "
import sys, json
try:
  d = json.loads(sys.stdin.read())
  print('G=$1 s3=$2 wbfm', d['value'], 'Msps frac', d['roofline']['frac'], 'ms', d['ms_per_step'])
except Exception as e: print('G=$1 s3=$2 failed')
" | tee -a gpurun_out/sweep_wb.txt
done
