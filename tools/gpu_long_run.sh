#!/bin/bash
# does a long timed region cost throughput (power cap, nvidia-smi polling, host run-ahead)?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  env $1 timeout 300 python bench.py --workload ${3:-am} --steps $2 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1 ${3:-am} steps $2:', d['value'], 'Msps  ms/step', d['ms_per_step'], (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'))
" | tee -a gpurun_out/long_run.txt
}
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
for pace in $@; do
  run "SDR_PACE=$pace SDR_BENCH_NO_SAMPLER=1" 15000
  run "SDR_PACE=$pace" 15000
done
