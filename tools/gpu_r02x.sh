#!/bin/bash
# timing experiment (results are wrong on purpose): the AM / SSB FIR kernel without stage 1's arithmetic
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload $WL --steps 200 --warmup 10 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for WL in am ssb; do
run A=1
run SDR_B200_LIB=$PWD/rtlsdrdiags_b200/libsdr_ab_fake1.so
run SDR_B200_LIB=$PWD/rtlsdrdiags_b200/libsdr_ab_fake2.so
done
