#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for sp in 0 1; do
for wl in am; do
  SDR_AM_SPLIT=$sp timeout 300 python bench.py --workload $wl --no-extras --no-cpu --steps 300 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('split=$sp $wl', d['value'], 'Msps frac', d['roofline']['frac'], 'ms', d['ms_per_step'])
" | tee -a gpurun_out/ab_am.txt
done; done
