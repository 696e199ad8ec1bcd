#!/bin/bash
# mixed bank: finer sweep of the minimum share lengths
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 100 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'])"; }
WL=mixed
run A=1
run SDR_FM_MIN_SHARE=4
run SDR_FM_MIN_SHARE=6
run SDR_AM_MIN_SHARE=12
run SDR_AM_MIN_SHARE=20
run A=1
