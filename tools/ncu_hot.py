#!/usr/bin/env python
"""Where a kernel's issue slots go, from an `ncu --set full --import-source on` report:
    python tools/ncu_hot.py rep.ncu-rep [top]
prints the executed-instruction mix by SASS opcode and the hottest source lines (file:line, warp
instructions executed, share). Runs here, no GPU needed."""
import subprocess, csv, io, sys, collections, re
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
def page(kind):
    r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", kind], capture_output=True, text=True)
    return r.stdout
# SASS: opcode mix
rows = list(csv.reader(io.StringIO(page("sass"))))
hdr = None; mix = collections.Counter(); total = 0
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    try: n = int(r[hdr.index("Instructions Executed")])
    except ValueError: continue
    src = r[hdr.index("Source")].strip()
    src = re.sub(r"^@!?U?P\d+\s+", "", src)
    op = src.split()[0] if src else "?"
    mix[op.split(".")[0]] += n; total += n
print("total warp instructions executed:", total)
for op, n in mix.most_common(top):
    print("  %-12s %12d  %5.1f%%" % (op, n, 100.0 * n / total))
# CUDA source lines with the SASS attributed to them: hottest lines
out = page("cuda,sass")
cur = None; hdr = None; lines = []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = r[1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; col = [i for i, h in enumerate(r) if h == "Instructions Executed"][0]; continue
    if hdr is None or not r[0]: continue
    try: n = int(r[col])
    except (ValueError, IndexError): continue
    if n: lines.append((n, cur.split("/")[-1], r[0], r[1].strip()[:100]))
lines.sort(reverse=True)
print("hottest source lines (warp instructions attributed to the line, inlined callees at their own lines):")
for n, f, ln, s_ in lines[:top]:
    print("  %10d %5.1f%%  %s:%s  %s" % (n, 100.0 * n / total, f, ln, s_))
