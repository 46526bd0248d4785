#!/usr/bin/env python
"""Static SASS statistics of the kernels one time step launches (fast build, unfused pipeline):
    python profiles/count_sass.py > profiles/sass_counts.json
Every kernel processes one cell / face / edge per thread, so the per-thread static count is the per-cell count
(both sides of the few data-dependent branches are included, so dynamic counts are slightly lower)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "ppkmhd_b200", "_build", "kernels_fast.o")
STEP = {  # kernel-name regex -> launches per step
    r"k_prim_dt": 1, r"k_elec_dbf": 1, r"k_traceILi4": 1,
    r"k_flux_tmaILi0ELb0ELi4": 1, r"k_flux_tmaILi1ELb0ELi4": 1, r"k_flux_tmaILi2ELb0ELi4": 1,
    r"k_emf_tmaILi0ELb0": 1, r"k_emf_tmaILi1ELb0": 1, r"k_emf_tmaILi2ELb0": 1, r"8k_updateILi0": 1,
}
sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
stats, name = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        stats[name] = {"instructions": 0, "fp64_pipe": 0, "mufu": 0, "ldg": 0, "lds": 0, "stg": 0, "utmaldg": 0}
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or name is None:
        continue
    op = m.group(1).split(".")[0]
    st = stats[name]
    st["instructions"] += 1
    st["fp64_pipe"] += op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
    st["mufu"] += op == "MUFU"
    st["ldg"] += op == "LDG"
    st["lds"] += op == "LDS"
    st["stg"] += op == "STG"
    st["utmaldg"] += op == "UTMALDG"
out, tot = {}, {"instructions": 0, "fp64_pipe": 0, "mufu": 0}
for pat, n in STEP.items():
    for k, v in stats.items():
        if re.search(pat, k):
            out[pat] = v
            for key in tot:
                tot[key] += n * v[key]
            break
    else:
        sys.exit(f"kernel {pat} not found in {OBJ}")
json.dump({"per_kernel_per_thread": out, "per_cell_update": tot,
           "note": "fast build (-DPPK_EXACT=0), unfused pipeline; TMA kernels include ~250 instructions of descriptor/"
                   "TMA issue code that only one warp per CTA executes"}, sys.stdout, indent=1)
