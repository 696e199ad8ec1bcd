#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02g_pytest.txt
for u in 1 2 3 4; do
lib=rtlsdrdiags_b200/libsdr_b200_unroll$u.so; [ $u = 2 ] && lib=""
SDR_B200_LIB=$lib timeout 300 python bench.py --steps 1000 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('am unroll $u steps1000', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], d['parity']['gpu_pcm_identical'])"
SDR_B200_LIB=$lib timeout 300 python bench.py --workload ssb --steps 300 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('ssb unroll $u', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])"
done | tee gpurun_out/r02g_unroll.txt
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('am steps20 run $i', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'])"; done | tee gpurun_out/r02g_steps20.txt
for g in 24 26 28; do
SDR_WB_G=$g timeout 300 python bench.py --workload wbfm --steps 100 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('wbfm generation 3, $g channels per CTA', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])"
done | tee gpurun_out/r02g_wbfm.txt
for wl in mixed fm; do timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['parity']['gpu_pcm_identical'])"; done | tee -a gpurun_out/r02g_wbfm.txt
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k "regex:amssb_fir|dc_block" -s 8 -c 6 --csv --log-file gpurun_out/r02g_am_launches.csv python bench.py --workload am --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > /dev/null 2>&1
grep -E "fir_kernel|dc_block" gpurun_out/r02g_am_launches.csv | cut -d, -f5,13- | tail -12
