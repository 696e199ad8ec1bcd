#!/bin/bash
# eight GPUs, BASELINE config 5 (mixed x65536) with the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/r03g_bench8.err | tail -1 > gpurun_out/r03g_bench8.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r03g_bench8.json'))
    print('N=8', d['config']['workload'], d['config']['channels_total'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac_box', d['roofline']['frac_of_box'], 'e2e', d['e2e']['value'], 'parity', d['parity']['per_rank'], 'am_weak', d['am_weak'])
except Exception as ex:
    print('bench N=8 failed', ex); print(open('gpurun_out/r03g_bench8.err').read()[-3000:])
PY
