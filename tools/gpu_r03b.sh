#!/bin/bash
# AM / SSB: stage 3's and the Hilbert transformer's windows through shared-memory rings (80 registers: six CTAs per SM fit)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 200 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_full_size.py tests/test_gpu_tma.py tests/test_gpu_recurrence.py -x -q 2>&1 | tail -3
for WL in am ssb; do run A=1; run SDR_AM_WARPS_PER_SM=24; run SDR_AM_WARPS_PER_SM=72; done
for wl in am ssb; do
timeout 300 ncu --set full --clock-control none -k "regex:amssb_fir" -s 4 -c 1 -f -o gpurun_out/prof_${wl}_r03b python bench.py --workload $wl --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > /dev/null 2>&1
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('driver-style', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks'])"
