#!/bin/bash
# the +-pi wrap as two FMAs (three instructions instead of five): parity, then the modes that use it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 900 python -m pytest tests/test_gpu_wbfm.py tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
WL=wbfm; run A=1; run SDR_WB_KERNEL=3
WL=fm; run A=1
WL=mixed; run A=1
