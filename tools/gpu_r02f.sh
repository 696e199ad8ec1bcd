#!/bin/bash
# WBFM generation 3 (two channels per worker warp): parity, then WBFM / mixed throughput for generations 2 and 3;
# AM with the 2-tile unrolled build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wbfm.py tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_squelch.py -x -q 2>&1 | tail -8
for gen in 2 3; do for wl in wbfm mixed; do
SDR_WB_KERNEL=$gen timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl generation $gen', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])"
done; done | tee gpurun_out/r02f_wbfm.txt
for lib in "" rtlsdrdiags_b200/libsdr_b200_unroll2.so; do
SDR_B200_LIB=$lib timeout 300 python bench.py --steps 1000 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('am lib=[$lib] steps1000', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], d['parity']['gpu_pcm_identical'])"
SDR_B200_LIB=$lib timeout 300 python bench.py --workload ssb --steps 300 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('ssb lib=[$lib]', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])"
done | tee gpurun_out/r02f_unroll.txt
SDR_WB_KERNEL=3 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wbfm_tile3" -s 2 -c 1 -f -o gpurun_out/prof_wbfm3_r02f python bench.py --workload wbfm --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_wbfm3.log 2>&1
