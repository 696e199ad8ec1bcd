#!/bin/bash
cd "$(dirname "$0")/.."
for ns in 2 3 4 5; do
  SDR_AM_NSEG=$ns timeout 300 python bench.py --workload am --no-extras --no-cpu --steps 400 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('nseg=$ns am', d['value'], 'Msps frac', d['roofline']['frac'], 'ms', d['ms_per_step'])
"
done
