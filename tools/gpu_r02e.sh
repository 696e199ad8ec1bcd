#!/bin/bash
# two GPUs: BASELINE config 4 through bench.py under torchrun (both arms), the multi-device bank (pytest, C++ driver),
# the host->device copy probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02e_topology.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_bank.py tests/test_gpu_tma.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r02e_bench2.err | tail -1 > gpurun_out/r02e_bench2.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02e_bench2.json'))
    print('N=2', d['config']['workload'], d['config']['channels_total'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac_box', d['roofline']['frac_of_box'], 'e2e', d['e2e']['value'], 'parity', d['parity']['per_rank'], 'am_weak', d['am_weak'])
except Exception as ex:
    print('bench N=2 failed', ex); print(open('gpurun_out/r02e_bench2.err').read()[-3000:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
for g in 1 2; do timeout 120 ./rtlsdrdiags_b200/b200_bank -g $g -n 16384 -m ssb -t 30 -b 32768; done 2>&1 | tee gpurun_out/r02e_bank.txt
timeout 120 ./tools/h2d_probe 512 20 2>&1 | tee gpurun_out/r02e_h2d.txt
for ld in 0 2 10 12; do
SDR_BENCH_TILE_LOADER=$ld timeout 300 python bench.py --steps 1000 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('loader $ld am steps1000', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done | tee gpurun_out/r02e_loaders.txt
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('am steps20 run $i', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'])"; done | tee -a gpurun_out/r02e_loaders.txt
