#!/bin/bash
# round-end validation on one GPU: smoke, parity tests, sanitizers, bench of every workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
bash tools/gpu_sanitize.sh 2>&1 | tee gpurun_out/sanitize_summary.txt
timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_reference.json | cut -c1-300
