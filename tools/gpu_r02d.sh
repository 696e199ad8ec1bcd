#!/bin/bash
# AM FIR kernel variants: duration, executed warp-instructions and issue utilisation per launch (ncu, kernel alone),
# then the pipelined step (bench) for the same variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tma.py tests/test_gpu_recurrence.py -x -q 2>&1 | tail -3
for ld in 0 16 12 10 28 26 4 2; do
SDR_BENCH_TILE_LOADER=$ld timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k "regex:amssb_fir" -s 4 -c 3 --csv --log-file gpurun_out/r02d_$ld.csv python bench.py --workload am --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02d_$ld.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
d={}
for r in rows[start+1:]:
    if len(r)>vi: d.setdefault(r[mi],[]).append(float(r[vi].replace(',','')))
print('loader $ld', rows[start+1][ki][:44], ' '.join('%s=%.1f'%(k.split('.')[0][-18:], sum(v)/len(v)) for k,v in d.items()), 'instr/tile=%.1f'%(sum(d['smsp__inst_executed.sum'])/len(d['smsp__inst_executed.sum'])/262144))
PY
done 2>&1 | tee gpurun_out/r02d_variants.txt
for ld in 0 16 12 28 4; do
SDR_BENCH_TILE_LOADER=$ld timeout 300 python bench.py --steps 1000 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('loader $ld am steps1000', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" | tee -a gpurun_out/r02d_variants.txt
done
