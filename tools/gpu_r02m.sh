#!/bin/bash
# WBFM pre-filter on the tensor cores: parity first (the WBFM tests, racecheck on a small case), then A/B of the
# WBFM and the mixed bank with SDR_WB_MMA=0 / 1, then ncu of the new kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wbfm.py -x -q 2>&1 | tail -15
for wl in wbfm mixed; do for m in 0 1; do
  echo "== $wl SDR_WB_MMA=$m"
  SDR_WB_MMA=$m timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['gpu_pcm_identical'], d['clocks'])"
done; done 2>&1 | tee gpurun_out/r02m_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wbfm_tile3" -s 4 -c 1 -f -o gpurun_out/prof_wbfm_r02m \
   python bench.py --workload wbfm --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_wbfm_r02m.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_wbfm.py::test_tensor_core_prefilter[2-2]" "tests/test_gpu_wbfm.py::test_tensor_core_prefilter[2-3]" -x -q 2>&1 | tail -8
