#!/usr/bin/env python
"""Condenses gpurun_out/*.ncu-rep + launch lists into small tracked files under profiles/.

    python tools/ncu_summary.py <tag> [workload ...]

For each workload it writes profiles/<tag>_<workload>_ncu.txt (selected raw metrics of the
`ncu --set full` capture, per-phase stall attribution from the source page) and copies the
launch list; profiles/ncu_traffic.json gets the per-launch DRAM traffic that bench.py
reports as roofline.traffic.
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg",
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "mio_throttle",
          "lg_throttle", "not_selected", "selected", "branch_resolving", "no_instruction", "dispatch_stall"]


def ncu_csv(rep, page):
    r = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


def summarise(tag, wl):
    rep = os.path.join(OUT, "prof_%s_%s.ncu-rep" % (wl, tag))
    if not os.path.exists(rep):
        print("missing", rep)
        return None
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    lines = ["ncu --set full --clock-control none, one launch of %s" % m.get("Kernel Name", ("?",))[0],
             "command: python bench.py --workload %s --steps 3 --warmup 3 --no-extras --no-cpu" % wl, ""]
    for k in WANT:
        if k in m:
            lines.append("%-72s %-16s %s" % (k, m[k][1], m[k][0]))
    lines.append("")
    lines.append("warps stalled per issued instruction, by reason (smsp__average_warps_issue_stalled_*_per_issue_active):")
    for s in STALLS:
        k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
        if k in m:
            lines.append("  %-24s %s" % (s, m[k][0]))
    # per-phase attribution: segments of the SASS between CTA barriers
    src = ncu_csv(rep, "source")
    if len(src) > 3:
        h = src[1]
        ci = {name: i for i, name in enumerate(h)}
        data = src[2:]
        tot = sum(int(r[ci["# Samples"]]) for r in data) or 1
        lines += ["", "stall samples and executed warp-instructions per code segment (segments end at BAR.SYNC;",
                  "a warp waiting at a barrier is sampled at the first instruction of the NEXT segment):"]
        seg, cur, inst = 0, 0, 0
        for r in data:
            cur += int(r[ci["# Samples"]])
            inst += int(r[ci["Instructions Executed"]])
            if "BAR.SYNC" in r[ci["Source"]]:
                lines.append("  segment %2d: %5.1f%% of samples, %11d warp-instructions" % (seg, 100.0 * cur / tot, inst))
                seg, cur, inst = seg + 1, 0, 0
        lines.append("  segment %2d: %5.1f%% of samples, %11d warp-instructions" % (seg, 100.0 * cur / tot, inst))
    with open(os.path.join(PROF, "%s_%s_ncu.txt" % (tag, wl)), "w") as f:
        f.write("\n".join(lines) + "\n")
    ll = os.path.join(OUT, "launches_%s_%s.csv" % (wl, tag))
    if os.path.exists(ll):
        shutil.copy(ll, os.path.join(PROF, "%s_%s_launches.csv" % (tag, wl)))

    def mb(key):
        v, u = m[key]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")


def main():
    tag = sys.argv[1]
    wls = sys.argv[2:] or ["am", "fm", "wbfm"]
    path = os.path.join(PROF, "ncu_traffic.json")
    traffic = json.load(open(path)) if os.path.exists(path) else {}
    for wl in wls:
        t = summarise(tag, wl)
        if t is not None:
            traffic[wl] = round(t)
    traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), from the ncu --set full "
                        "captures named in profiles/*_ncu.txt")
    json.dump(traffic, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
