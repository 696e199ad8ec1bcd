#!/bin/bash
# defaults after the policy change (generation 4 for large banks with the wrap vote, packed generation 2 in mixed banks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 600 python -m pytest tests/test_gpu_wbfm.py -x -q 2>&1 | tail -3
WL=wbfm; run A=1; run SDR_WB_KERNEL=3
WL="wbfm --signal noise"; run A=1; run SDR_WB_KERNEL=3
WL=mixed; run A=1; run SDR_WB_GX=15; run SDR_WB_GX=12
