#!/bin/bash
# usage: tools/gpu_prof_kernel.sh <workload> <kernel-regex> <tag>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -f -o gpurun_out/prof_$3 \
   python bench.py --workload $1 --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
