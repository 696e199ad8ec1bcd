#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 600 python -m pytest tests/test_gpu_wbfm.py -x -q  2>&1 | grep -E "assert|Error|passed|failed" | head -8
WL=wbfm; run SDR_WB_KERNEL=4; run A=1; WL="wbfm --signal noise"; run A=1; WL=mixed; run A=1
