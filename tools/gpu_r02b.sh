#!/bin/bash
# AM step: launch timeline for both tile loaders and several recurrence segmentations; per-kernel durations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tma.py tests/test_gpu_recurrence.py tests/test_gpu_bank.py tests/test_gpu_parity.py tests/test_gpu_ingest.py -x -q 2>&1 | tail -5
SDR_TIMELINE_ROWS=1 timeout 600 python tools/probe_timeline.py sweep 2>&1 | tee gpurun_out/r02b_timeline.txt | grep -v "call "
for ld in 4 12 0; do
SDR_BENCH_TILE_LOADER=$ld timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fir_kernel|dc_block" -c 40 --csv --log-file gpurun_out/r02b_am_launches_$ld.csv python bench.py --steps 10 --warmup 3 --no-extras --no-cpu --no-e2e > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02b_am_launches_$ld.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[start+1:]:
    if len(r)>vi: d[r[ki][:60]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print('$ld', k, len(v), 'avg %.1f us'%(sum(v)/len(v)/1000), 'min %.1f'%(min(v)/1000),'max %.1f'%(max(v)/1000))
PY
done
