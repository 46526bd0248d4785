#!/usr/bin/env python
"""Per-instruction stall profile of one kernel from an ncu report (SASS view): prints the hottest instructions and
totals per stall reason. Usage: python profiles/srcprof.py X.ncu-rep [top]"""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first row: kernel name; second: header
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}; total = 0
for r in data:
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
    total += int(r[col["# Samples"]] or 0)
print("kernel:", rows[0][1][:90]); print("samples", total, "instructions (SASS lines)", len(data))
print("executed warp-instr:", sum(int(r[col["Instructions Executed"]] or 0) for r in data))
print({k.replace("stall_", ""): round(100 * v / max(total, 1), 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = data[i]
    st = sorted(((int(r[col[s]] or 0), s.replace("stall_", "")) for s in stalls), reverse=True)[:3]
    print(f"{i:5d} {int(r[col['# Samples']]):6d} {r[col['Instructions Executed']]:>9s}  {r[col['Source']].strip()[:70]:70s} {st}")
