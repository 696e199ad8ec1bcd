#!/bin/bash
# WBFM pre-filter on FFMA2 (default build) against IDP.2A (libsdr_b200_wbidp.so): parity, then WBFM x8192 and mixed x8192
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wbfm.py tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_vs_reference.py tests/test_capture.py -x -q 2>&1 | tail -4
for lib in "" rtlsdrdiags_b200/libsdr_b200_wbidp.so; do for wl in wbfm mixed; do for sig in tone noise; do
SDR_B200_LIB=$lib timeout 300 python bench.py --workload $wl --signal $sig --steps 100 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl $sig lib=[$lib]', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])"
done; done; done | tee gpurun_out/r02k_wbfm_fp32.txt
SDR_WB_KERNEL=2 timeout 300 python bench.py --workload wbfm --steps 100 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('wbfm generation 2 with the FP32 pre-filter', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])" | tee -a gpurun_out/r02k_wbfm_fp32.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wbfm_tile3" -s 2 -c 1 -f -o gpurun_out/prof_wbfm3_r02k python bench.py --workload wbfm --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_wbfm3_r02k.log 2>&1
