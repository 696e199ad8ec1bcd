#!/bin/bash
# AM / SSB FIR kernel at 80 registers (six resident CTAs per SM): how many shares per SM?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 200 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for rep in 1 2; do
WL=am; for w in 48 72 96; do run SDR_AM_WARPS_PER_SM=$w; done
done
WL=ssb; for w in 48 72; do run SDR_AM_WARPS_PER_SM=$w; done
for w in 48 72; do
echo "== driver-style AM $w"; SDR_AM_WARPS_PER_SM=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
