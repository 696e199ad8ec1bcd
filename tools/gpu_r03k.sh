#!/bin/bash
# small shares of a mixed bank, AM / SSB and NBFM: minimum share length
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 100 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
WL=mixed
for a in 16 32; do for f in 8 16 32; do run SDR_AM_MIN_SHARE=$a SDR_FM_MIN_SHARE=$f; done; done
run SDR_AM_MIN_SHARE=20 SDR_FM_MIN_SHARE=16
