#!/bin/bash
# generation 4 of the WBFM kernel (tcgen05 pre-filter): parity, then A/B against generation 3 / 2, then ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wbfm.py -x -q 2>&1 | tail -15
for wl in wbfm mixed; do for g in 3 4; do
  echo "== $wl SDR_WB_KERNEL=$g"
  SDR_WB_KERNEL=$g timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done 2>&1 | tee gpurun_out/r02n_ab.txt
SDR_WB_KERNEL=4 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wbfm_tile4" -s 4 -c 1 -f -o gpurun_out/prof_wbfm_r02n \
   python bench.py --workload wbfm --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_wbfm_r02n.log 2>&1
tail -2 gpurun_out/ncu_wbfm_r02n.log
