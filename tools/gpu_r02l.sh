#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in rtlsdrdiags_b200/libsdr_b200_wbfp1.so rtlsdrdiags_b200/libsdr_b200_wbfp2.so rtlsdrdiags_b200/libsdr_b200_wbfp4.so rtlsdrdiags_b200/libsdr_b200_wbidp.so; do for wl in wbfm mixed; do
SDR_B200_LIB=$lib timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl lib=[$lib]', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])"
done; done | tee gpurun_out/r02l_wbfm_fp32_split.txt
