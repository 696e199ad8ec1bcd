#!/usr/bin/env python
"""Selected raw metrics and per-issue stall reasons of `ncu --set full` captures, one block per report:
    python tools/ncu_brief.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/<name>.txt
Runs here (no GPU needed): `ncu -i <rep> --page raw --csv`."""
import subprocess, csv, io, sys
WANT = ["gpu__time_duration.sum","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active",
"dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"sm__inst_issued.avg.pct_of_peak_sustained_active","sm__inst_executed.sum","smsp__inst_executed.sum",
"sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
"sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tensor.sum","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","sm__cycles_elapsed.avg","sm__inst_executed_pipe_uniform.sum","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct"]
STALLS = ["barrier","long_scoreboard","short_scoreboard","wait","math_pipe_throttle","mio_throttle","lg_throttle","not_selected","selected","branch_resolving","no_instruction","dispatch_stall","tex_throttle","sleeping","membar","drain","imc_miss"]
for rep in sys.argv[1:]:
    r = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True)
    rows=list(csv.reader(io.StringIO(r.stdout)))
    hdr,units,vals=rows[0],rows[1],rows[2]
    m={h:(vals[i],units[i]) for i,h in enumerate(hdr)}
    print("==",rep, m.get("Kernel Name",("?",))[0][:70])
    for k in WANT:
        if k in m: print("  %-70s %-14s %s"%(k,m[k][1],m[k][0]))
    for s in STALLS:
        k="smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"%s
        if k in m: print("  stall %-20s %s"%(s,m[k][0]))
    for k in m:
        if 'pipe' in k and 'pct_of_peak_sustained_active' in k and k not in WANT:
            try:
                if float(m[k][0])>5: print("  ",k,m[k][0])
            except: pass
