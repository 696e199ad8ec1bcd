#!/bin/bash
# AM at 80 registers: does a deeper TMA pipeline pay now that six CTAs are resident anyway?
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 200 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
WL=am; run A=1; run SDR_BENCH_TILE_LOADER=3; run SDR_BENCH_TILE_LOADER=4; run SDR_BENCH_TILE_LOADER=0; run A=1
WL=ssb; run A=1; run SDR_BENCH_TILE_LOADER=3
