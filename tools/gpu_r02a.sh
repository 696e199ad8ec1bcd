#!/bin/bash
# round 2, first call: parity of the segment-parallel recurrence + AM step spread + box topology
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( nvidia-smi topo -m; lscpu | grep -iE "model name|socket|numa|^cpu\(s\)"; free -g | head -2;
  for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [ "$c" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) $(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done ) > gpurun_out/r02_topology.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02a_pytest.txt
for i in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('am steps20 run $i', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks']['sm_mhz'])" | tee -a gpurun_out/r02a_am.txt
done
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('am steps2000', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks'])" | tee -a gpurun_out/r02a_am.txt
SDR_BENCH_TILE_LOADER=cpasync timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('am steps2000 cp.async loader', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], d['clocks'])" | tee -a gpurun_out/r02a_am.txt
for wl in ssb mixed; do
timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'])" | tee -a gpurun_out/r02a_am.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02a_am_launches.csv python bench.py --steps 10 --warmup 3 --no-extras --no-cpu --no-e2e > /dev/null 2>&1
grep -E "fir_kernel|dc_block" gpurun_out/r02a_am_launches.csv | tail -8
