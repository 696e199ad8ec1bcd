#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 20 50 100 200 500 2000; do
  env $1 timeout 300 python bench.py --workload am --steps $k --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1 steps $k:', d['value'], 'Msps  ms/step', d['ms_per_step'])
" | tee -a gpurun_out/short_runs.txt
done
