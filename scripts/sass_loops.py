#!/usr/bin/env python
"""Instruction mix of the innermost loops of one kernel of libvr180_b200.so (design aid).
    python scripts/sass_loops.py 'LinearELb0ELi2' [n_loops] [--dump]"""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
pat = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "vr180-convert_b200" / "libvr180_b200.so")], capture_output=True, text=True).stdout
funcs, cur = {}, None
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name: continue
    a2i = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in a2i:
                body = ins[a2i[tgt]:i + 1]
                nl = sum(1 for _, x in body if re.search(r"\bLDS\b", x))
                if nl >= 8: loops.append((len(body), nl, tgt, a))
    loops.sort()
    print(name, "total", len(ins))
    for ln, nl, tgt, a in loops[:n]:
        body = ins[a2i[tgt]:a2i[a] + 1]
        c = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for _, t in body)
        print(" loop", hex(tgt), hex(a), "len", ln, "LDS", nl, c.most_common(16))
        if "--dump" in sys.argv:
            for a_, t in body: print("   ", hex(a_), t)
