#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_full_size.py tests/test_gpu_recurrence.py tests/test_gpu_bank.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload mixed --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mixed', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
