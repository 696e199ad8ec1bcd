#!/bin/bash
# AM / SSB stage 1 with the front end grouped by operation (six instructions per rotation period instead of ten)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 200 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_full_size.py tests/test_gpu_tma.py tests/test_gpu_recurrence.py tests/test_gpu_squelch.py tests/test_gpu_ingest.py -x -q 2>&1 | tail -3
WL=am; run A=1; run A=2
WL=ssb; run A=1
WL=mixed; run A=1
for wl in am ssb; do
timeout 300 ncu --set full --clock-control none -k "regex:amssb_fir" -s 4 -c 1 -f -o gpurun_out/prof_${wl}_r03e python bench.py --workload $wl --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > /dev/null 2>&1
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('driver-style', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'])"
