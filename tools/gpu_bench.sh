#!/bin/bash
# bench + launch list + one full ncu capture of the headline kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.txt
timeout 900 python bench.py "$@" 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
