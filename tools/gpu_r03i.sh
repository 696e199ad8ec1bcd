#!/bin/bash
# two GPUs, BASELINE config 4 (LSB/USB x16384) with the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu 2>gpurun_out/r03i_bench4.err | tail -1 > gpurun_out/r03i_bench4.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r03i_bench4.json'))
    print('N=4', d['config']['workload'], d['config']['channels_total'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac_box', d['roofline']['frac_of_box'], 'e2e', d['e2e']['value'], 'parity', d['parity']['per_rank'], 'am_weak', d['am_weak'])
except Exception as ex:
    print('bench N=4 failed', ex); print(open('gpurun_out/r03i_bench4.err').read()[-3000:])
PY
