#!/bin/bash
# ncu --set full of the AM FIR kernel for three loaders: cp.async (0), TMA x4 SIMT (12), TMA x4 + tensor-core stage 1 (4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ld in 0 12 4; do
SDR_BENCH_TILE_LOADER=$ld timeout 600 ncu --set full --clock-control none --import-source on -k "regex:amssb_fir" -s 4 -c 1 -f -o gpurun_out/prof_am_r02c_$ld \
   python bench.py --workload am --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_am_r02c_$ld.log 2>&1
tail -1 gpurun_out/ncu_am_r02c_$ld.log | cut -c1-200
done
SDR_BENCH_TILE_LOADER=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:dc_block" -s 4 -c 1 -f -o gpurun_out/prof_dc_r02c \
   python bench.py --workload am --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_dc_r02c.log 2>&1
ls -la gpurun_out/*.ncu-rep
