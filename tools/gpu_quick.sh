#!/bin/bash
# parity tests + a short bench of the given workloads; everything under `timeout`
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
for wl in "$@"; do
  timeout 300 python bench.py --workload $wl --no-extras --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$wl', d['value'], 'Msps  frac', d['roofline']['frac'], ' ms', d['ms_per_step'], ' e2e', d['e2e']['value'], 'sync', d['e2e'].get('synchronous_value'), d['e2e'].get('paths_agree'))
except Exception as ex:
    print('$wl bench failed', ex)
" | tee -a gpurun_out/quick_bench.txt
done
