#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed):  python profiles/summarize.py gpurun_out/X.ncu-rep > profiles/X.md
Prints, per captured launch: duration, DRAM bytes read/written, DRAM %, FP64-pipe %, issue %, registers, occupancy,
L1/L2 hit rates and the top warp-stall reasons (pc sampling)."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    print(f"# ncu summary of `{path}` ({len(data)} launches; --set full --clock-control none)\n")
    print("| kernel | " + " | ".join(n for _, n in WANT) + " | top stalls (pc samples) |")
    print("|---|" + "---|" * (len(WANT) + 1))
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for key, _ in WANT:
            if key in col:
                v, u = r[col[key]], units[col[key]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("n/a")
        st = sorted(((float(r[col[s]].replace(",", "") or 0), s.replace("smsp__pcsamp_warps_issue_stalled_", "")) for s in stalls), reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:4])
        print(f"| {name} | " + " | ".join(cells) + f" | {top} |")


if __name__ == "__main__":
    main(sys.argv[1])
