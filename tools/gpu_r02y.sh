#!/bin/bash
# generation 4 in a long run and through the C++ bank driver (ingest ring, pinned host buffers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --workload wbfm --steps 1500 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('wbfm 1500 steps', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity'], d['clocks'])"
timeout 200 ./rtlsdrdiags_b200/b200_bank -g 1 -n 8192 -m wbfm -t 40 -b 32768
timeout 200 ./rtlsdrdiags_b200/b200_bank -g 1 -n 8192 -m mixed -t 40 -b 32768
timeout 600 python bench.py --workload wbfm --steps 20 --warmup 5 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('wbfm with e2e + cpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('paths_agree'), d['cpu_baseline']['value'], d['cpu_baseline'].get('gpu_pcm_identical'))"
