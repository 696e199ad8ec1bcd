#!/bin/bash
# final evidence of the round on one GPU: ncu --set full of each workload's dominant kernel (DRAM traffic for
# roofline.traffic, instruction counts, stalls), launch lists, and the driver's own bench commands
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
mkdir -p gpurun_out
for spec in "am:amssb_fir" "ssb:amssb_fir" "fm:fm_tile" "wbfm:wbfm_tile4" "mixed:wbfm_tile2"; do
wl=${spec%%:*}; k=${spec##*:}
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 4 -c 1 -f -o gpurun_out/prof_${wl}_r03f \
   python bench.py --workload $wl --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ncu_${wl}_r03f.log 2>&1
done
for wl in am mixed; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fir_kernel|dc_block|tile_kernel|tile2_kernel|tile3_kernel|tile4_kernel" -s 12 -c 60 --csv --log-file gpurun_out/r03f_${wl}_launches.csv python bench.py --workload $wl --steps 10 --warmup 3 --no-extras --no-cpu --no-e2e > /dev/null 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r03f_bench.json 2> gpurun_out/r03f_bench.err
tail -c 600 gpurun_out/r03f_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r03f_bench_reference.json 2>&1
tail -c 300 gpurun_out/r03f_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
