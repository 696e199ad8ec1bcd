#!/bin/bash
# WBFM kernel A/B: parity tests with the new kernel, then wbfm / mixed bench with both generations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
for gen in 1 2; do
  for wl in wbfm mixed; do
    SDR_WB_KERNEL=$gen timeout 300 python bench.py --workload $wl --no-extras --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('gen $gen $wl', d['value'], 'Msps  frac', d['roofline']['frac'], ' ms', d['ms_per_step'])
except Exception as ex:
    print('gen $gen $wl bench failed', ex)
" | tee -a gpurun_out/wb_ab.txt
  done
done
for g in $@; do
  SDR_WB_G=$g timeout 300 python bench.py --workload wbfm --no-extras --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('gen 2 G=$g wbfm', d['value'], 'Msps  frac', d['roofline']['frac'], ' ms', d['ms_per_step'])
except Exception as ex:
    print('G=$g bench failed', ex)
" | tee -a gpurun_out/wb_ab.txt
done
