#!/bin/bash
# eight GPUs: BASELINE config 5 (mixed x65536) through bench.py under torchrun, both arms; the multi-device bank
# from C++ (one host thread) at 1/2/4/8 GPUs; the host->device copy probe; the bank tests over all devices
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02i_topology.txt 2>&1
lscpu | grep -iE "model name|socket|numa|^cpu\(s\)" >> gpurun_out/r02i_topology.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/r02i_bench8.err | tail -1 > gpurun_out/r02i_bench8.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02i_bench8.json'))
    print('N=8', d['config']['workload'], d['config']['channels_total'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac_box', d['roofline']['frac_of_box'], 'e2e', d['e2e']['value'], 'parity', d['parity']['per_rank'], 'am_weak', d['am_weak'])
except Exception as ex:
    print('bench N=8 failed', ex); print(open('gpurun_out/r02i_bench8.err').read()[-3000:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
for g in 1 2 4 8; do timeout 200 ./rtlsdrdiags_b200/b200_bank -g $g -n 65536 -m mixed -t 20 -b 32768; done 2>&1 | tee gpurun_out/r02i_bank.txt
for g in 4 8; do timeout 200 ./rtlsdrdiags_b200/b200_bank -g $g -n 16384 -m ssb -t 40 -b 32768; done 2>&1 | tee -a gpurun_out/r02i_bank.txt
timeout 200 ./tools/h2d_probe 512 20 2>&1 | tee gpurun_out/r02i_h2d.txt
timeout 300 python -m pytest tests/test_gpu_bank.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu 2>gpurun_out/r02i_bench4.err | tail -1 > gpurun_out/r02i_bench4.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02i_bench4.json'))
    print('N=4', d['config']['workload'], d['config']['channels_total'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac_box', d['roofline']['frac_of_box'], 'e2e', d['e2e']['value'], 'parity', d['parity']['per_rank'], 'am_weak', d['am_weak'])
except Exception as ex:
    print('bench N=4 failed', ex); print(open('gpurun_out/r02i_bench4.err').read()[-3000:])
PY
