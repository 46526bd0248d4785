#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode of one kernel from an ncu report (source page).
Usage: python profiles/ophist.py X.ncu-rep [cells]   -> warp-instructions x32 / cells = thread-instructions per cell"""
import csv, io, subprocess, sys, collections
path = sys.argv[1]; cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
h = collections.Counter(); tot = 0
for r in data:
    src = r[col["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0] if not op.startswith(("LDS", "LDG", "STG", "STS", "MUFU", "F2F", "I2F")) else ".".join(op.split(".")[:2])
    n = int(r[col["Instructions Executed"]] or 0)
    h[op] += n; tot += n
print("total warp-inst", tot, ("= %.0f thread-inst/cell" % (tot * 32 / cells)) if cells else "")
for op, n in h.most_common(40):
    print(f"{op:16s} {n:12d} {100*n/tot:5.1f}%" + (f" {n*32/cells:8.1f}/cell" if cells else ""))
