#!/bin/bash
# generation 4 in the mixed bank: which geometry / channels per CTA; and the large WBFM bank's channels per CTA
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
WL=mixed
run SDR_WB_KERNEL=2
run SDR_WB_KERNEL=4 SDR_WB4_TWO=0
run SDR_WB_KERNEL=4 SDR_WB4_TWO=0 SDR_WB_GX=14
run SDR_WB_KERNEL=4 SDR_WB4_TWO=1
run SDR_WB_KERNEL=4 SDR_WB4_TWO=1 SDR_WB_GX=20
run SDR_WB_KERNEL=4 SDR_WB4_TWO=1 SDR_WB_GX=28
run SDR_WB_KERNEL=3 SDR_WB_G=28
WL=wbfm
run SDR_WB_KERNEL=4
run SDR_WB_KERNEL=4 SDR_WB_GX=26
run SDR_WB_KERNEL=4 SDR_WB_GX=24
run SDR_WB_KERNEL=4 SDR_WB4_TWO=0
run SDR_WB_KERNEL=2
