#!/bin/bash
# small AM / SSB shares of a mixed bank: how short may a share be before its warm-up tiles cost more than the parallelism buys?
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 100 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
WL=mixed; for m in 8 16 24 32 8; do run SDR_AM_MIN_SHARE=$m; done
