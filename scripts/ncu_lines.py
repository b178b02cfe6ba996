#!/usr/bin/env python
"""Per-source-line executed warp instructions and stall samples of an ncu report (needs --import-source on):
    python scripts/ncu_lines.py REPORT.ncu-rep [top]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
def num(x):
    try: return int(x)
    except ValueError: return 0
cur_file = ""; agg = defaultdict(lambda: [0, 0, ""]); hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != "":  # a source line summary row
        key = (cur_file, int(r[0]))
        agg[key][0] += num(r[hdr.index("Instructions Executed")])
        agg[key][1] += num(r[hdr.index("# Samples")])
        agg[key][2] = r[1].strip()[:90]
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot_i}  samples {tot_s}")
byfile = defaultdict(lambda: [0, 0])
for (f, l), v in agg.items(): byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]): print(f"  {f:20s} inst {100*v[0]/tot_i:5.1f}%  samples {100*v[1]/max(tot_s,1):5.1f}%")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l:<5d} inst {100*v[0]/tot_i:5.2f}%  samp {100*v[1]/max(tot_s,1):5.2f}%  {v[2]}")
