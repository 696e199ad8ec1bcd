#!/bin/bash
# generation 4 when every tile falls back (uniform random bytes): where does the time go?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wbfm_tile4" -s 4 -c 1 -f -o gpurun_out/prof_wbfm_noise_r02t \
   python bench.py --workload wbfm --signal noise --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_wbfm_noise_r02t.log 2>&1
SDR_WB_KERNEL=3 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wbfm_tile3" -s 4 -c 1 -f -o gpurun_out/prof_wbfm_noise3_r02t \
   python bench.py --workload wbfm --signal noise --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_wbfm_noise3_r02t.log 2>&1
tail -1 gpurun_out/ncu_wbfm_noise_r02t.log | cut -c1-200
