#!/bin/bash
cd "$(dirname "$0")/.."
wl=${1:-am}
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size --clock-control none -k "regex:tile_kernel|fir_kernel|dc_block" -c 8 --csv python bench.py --workload $wl --steps 3 --warmup 3 --no-extras --no-cpu 2>/dev/null | grep -E "fir_kernel|dc_block|tile_kernel" | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    print(r[4][:40].ljust(40), r[-3].ljust(55), r[-1])
" | tail -12
