#!/usr/bin/env python
"""3-D fixtures for the limiter settings the main set (tests/golden/, all slope_type=2) does not reach, written by RUNNING THE
UNMODIFIED REFERENCE exactly like tests/golden/make_golden.py:

    slope_type=1  minmod in slope_unsplit_hydro_3d / slope_unsplit_mhd_3d (MHDBaseFunctor3D.h:362-495, 561-668)
    slope_type=0  neither 1 nor 2: hydro slopes are zero (:463-493), face-field slopes are limited with min(slope_type, 2) = 0

    make -C oracle ref && python tests/golden_slope/make_golden_slope.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OT = "[OrszagTang]\nkt=1\n"
BLAST = "[blast]\nradius=0.25\ndensity_in=1.0\ndensity_out=1.2\npressure_in=10.0\npressure_out=0.1\n"
CASES = {
    # name: (problem, (nx,ny,nz), nsteps, extra, bc, slope_type)
    "ot_minmod_16x12x8": ("orszag_tang", (16, 12, 8), 5, OT, 3, "1"),
    "blast_mixedbc_minmod_12x10x8": ("blast", (12, 10, 8), 6, BLAST, [1, 2, 3, 3, 2, 1], "1"),
    "ot_noslope_16x12x8": ("orszag_tang", (16, 12, 8), 5, OT, 3, "0"),
}


def run(case):
    problem, n, nsteps, extra, bc, st = CASES[case]
    out = {}
    for tag, ns in (("step1", 1), ("stepN", nsteps)):
        ini = O.make_ini(problem=problem, n=n, nstepmax=ns, extra=extra, bc=bc, nlog=1, tend=10.0).replace("slope_type=2", "slope_type=" + st)
        stdout, states = O.run_reference(ini, threads=4)
        assert len(states) == 2, stdout
        out["init"] = states[0]
        out[tag] = states[1]
        if tag == "stepN":
            log = re.findall(r"time step=\s*(\d+) \(dt=\s*([-0-9.eE+]+) t=\s*([-0-9.eE+]+)\)", stdout)
            out["log_dt"] = np.array([float(m[1]) for m in log])
            out["final_time"] = np.array(float(re.search(r"final time is ([-0-9.eE+]+)", stdout).group(1)))
            out["ini"] = np.array(ini)
            out["nsteps"] = np.array(ns)
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for case in CASES:
        data = run(case)
        np.savez_compressed(os.path.join(here, case + ".npz"), **data)
        print(case, {k: getattr(v, "shape", None) for k, v in data.items()})
