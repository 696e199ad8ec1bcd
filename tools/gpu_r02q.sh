#!/bin/bash
# generation 4 tuning: pipelined tensor-memory loads / per-M-block barriers (A/B builds), channels per CTA in the mixed bank
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 600 python -m pytest tests/test_gpu_wbfm.py -x -q 2>&1 | tail -3
WL=wbfm
run SDR_WB_KERNEL=4
for v in p0b0 p1b0 p0b1; do run SDR_WB_KERNEL=4 SDR_B200_LIB=$PWD/rtlsdrdiags_b200/libsdr_ab_$v.so; done
run SDR_WB_KERNEL=4
WL=mixed
run SDR_WB_KERNEL=2 SDR_WB_GX=14
run SDR_WB_KERNEL=4 SDR_WB4_TWO=0 SDR_WB_GX=14
run SDR_WB_KERNEL=4 SDR_WB4_TWO=0 SDR_WB_GX=13
run SDR_WB_KERNEL=4 SDR_WB4_TWO=1 SDR_WB_GX=16
run SDR_WB_KERNEL=4 SDR_WB4_TWO=1 SDR_WB_GX=24
