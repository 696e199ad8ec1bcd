#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02h_pytest.txt
for lib in "" rtlsdrdiags_b200/libsdr_b200_fm2.so; do for sig in tone noise; do
SDR_B200_LIB=$lib timeout 300 python bench.py --workload fm --signal $sig --steps 100 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('fm $sig lib=[$lib]', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])"
done; done | tee gpurun_out/r02h_fm.txt
