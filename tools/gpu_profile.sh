#!/bin/bash
# ncu evidence for one workload: launch list (per-launch device time) + one full capture
# usage: tools/gpu_profile.sh <workload> [tag]
cd "$(dirname "$0")/.."
wl=${1:-am}; tag=${2:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --workload $wl --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tile_kernel|fir_kernel|dc_block" -c 12 \
    --csv --log-file gpurun_out/launches_${wl}_${tag}.csv $BENCH > gpurun_out/ncu_launches_${wl}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tile_kernel|fir_kernel|dc_block" -s 4 -c 1 \
    -f -o gpurun_out/prof_${wl}_${tag} $BENCH > gpurun_out/ncu_full_${wl}.log 2>&1
tail -3 gpurun_out/ncu_full_${wl}.log
ls -la gpurun_out/
