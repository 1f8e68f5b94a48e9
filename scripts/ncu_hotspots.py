"""Per-source-line / per-phase summary of an ncu report's source page (needs -lineinfo and --import-source on).

usage: python scripts/ncu_hotspots.py report.ncu-rep [kernel-substring] [phase-spec]
phase-spec: comma list of name:first_line for one file, e.g. "score_v2.cuh=step:54,staging:134,H:190"
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
kfilter = sys.argv[2] if len(sys.argv) > 2 else ""
phases = []
pfile = None
if len(sys.argv) > 3:
    pfile, spec = sys.argv[3].split("=")
    for item in spec.split(","):
        n, l = item.split(":")
        phases.append((int(l), n))
    phases.sort()
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, cur_line, cur_src, func = None, None, "", ""
agg = collections.OrderedDict()
ops = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        func = r[1]
        continue
    if not r or r[0] == "Line No":
        continue
    if r[0].isdigit():
        cur_line = int(r[0])
        cur_src = r[1]
        continue
    if r[0] == "" and len(r) > 8 and r[2].startswith("0x") and kfilter in func:
        try:
            samples, inst = int(r[6]), int(r[7])
        except ValueError:
            continue
        key = (cur_file, cur_line)
        a = agg.setdefault(key, [0, 0, cur_src])
        a[0] += inst
        a[1] += samples
        op = r[3].split()
        op = [o for o in op if not o.startswith("@")]
        if op:
            ops[(key, op[0].split(".")[0])] += inst
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"# kernel filter '{kfilter}': total warp instructions {ti}, stall samples {ts}")
if phases:
    pi, ps = collections.Counter(), collections.Counter()
    for (f, l), (i, s, _) in agg.items():
        name = f"other files"
        if f == pfile:
            name = "preamble"
            for first, n in phases:
                if l >= first:
                    name = n
        pi[name] += i
        ps[name] += s
    for k in pi:
        print(f"{k:24s} inst {100 * pi[k] / ti:5.1f}%  samples {100 * ps[k] / ts:5.1f}%")
print()
for (f, l), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    top = sorted(((o, n) for (k, o), n in ops.items() if k == (f, l)), key=lambda x: -x[1])[:4]
    print(f"{f:18s} L{l:4d} inst {100 * i / ti:5.1f}% samp {100 * s / ts:5.1f}%  {src.strip()[:70]:70s} {[(o, round(100 * n / ti, 1)) for o, n in top]}")
