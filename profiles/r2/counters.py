#!/usr/bin/env python
"""ncu CSV (profiles/r2/counters.sh) -> profiles/r2/counters.json entry: per kernel of one step, DRAM bytes and executed
instructions.  Usage: python profiles/r2/counters.py gpurun_out/r2_counters_<n>_<pipeline>.csv <n> <pipeline> <steps>"""
import csv, json, os, re, sys

KEY = [(r"k_plane_group", "flux_xy_emf_z"), (r"k_xz_group", "flux_z_emf_y"), (r"k_trace_tma", "trace"), (r"k_producer", "producer"), (r"k_riemann_all|k_riemann_pers", "riemann_all"), (r"k_prim_dt<0>", "dt_only"),
       (r"k_prim_dt", "prim_dt"), (r"k_elec_dbf", "elec_dbf"), (r"k_trace", "trace"),
       (r"k_flux(_tma)?<0", "flux_x"), (r"k_flux(_tma)?<1", "flux_y"), (r"k_flux(_tma)?<2", "flux_z"),
       (r"k_emf(_tma)?<2", "emf_z"), (r"k_emf(_tma)?<1", "emf_y"), (r"k_emf(_tma)?<0", "emf_x"),
       (r"k_update", "update"), (r"k_boundary", "boundary"), (r"k_finalize_dt", "finalize_dt"), (r"k_advance_time", "advance_time")]


def main(path, n, pipeline, steps):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[1:]:
        name, metric, unit, val = r[col["Kernel Name"]], r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
        key = next((k for pat, k in KEY if re.search(pat, name)), None)
        if key is None:
            continue
        e = out.setdefault(key, {"dram_bytes": 0.0, "inst_executed": 0.0, "inst_fp64": 0.0, "ms": 0.0, "launches": 0.0})
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if metric.startswith("dram__bytes"):
            e["dram_bytes"] += val * scale / steps
        elif metric == "smsp__inst_executed.sum":
            e["inst_executed"] += val / steps
            e["launches"] += 1.0 / steps
        elif metric == "smsp__inst_executed_pipe_fp64.sum":
            e["inst_fp64"] += val / steps
        elif metric == "gpu__time_duration.sum":
            e["ms"] += val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6) / steps
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "counters.json")
    allj = json.load(open(dst)) if os.path.exists(dst) else {}
    allj[f"{pipeline}_{n}"] = {"n": n, "cells": float(n) ** 3, "pipeline": pipeline, "steps_captured": steps,
                               "source": os.path.basename(path) + " (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,"
                                         "smsp__inst_executed_pipe_fp64.sum,gpu__time_duration.sum --clock-control none; per step)",
                               "kernels": out}
    json.dump(allj, open(dst, "w"), indent=1, sort_keys=True)
    tot = {k: sum(v[k] for v in out.values()) for k in ("dram_bytes", "inst_executed", "inst_fp64", "ms")}
    cells = float(n) ** 3
    print(f"{pipeline} {n}^3: {tot['dram_bytes'] / 1e9:.2f} GB DRAM/step = {tot['dram_bytes'] / cells:.0f} B/cell, {tot['inst_executed'] * 32 / cells:.0f} inst/cell, "
          f"{tot['inst_fp64'] * 32 / cells:.0f} FP64-pipe inst/cell, {tot['ms']:.2f} ms (serialised, under ncu)")
    for k, v in sorted(out.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"  {k:13s} {v['ms']:7.3f} ms  {v['dram_bytes'] / 1e9:6.2f} GB  {v['inst_executed'] * 32 / cells:7.0f} inst/cell  {v['inst_fp64'] * 32 / cells:6.0f} fp64/cell  x{v['launches']:.0f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 3)
