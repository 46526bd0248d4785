#!/usr/bin/env python
"""Golden fixtures of the 2-D MHD path (MHD_Muscl_2D, implementationVersion 0), written by RUNNING THE UNMODIFIED
REFERENCE (oracle/_ref/ppkMHD) exactly like tests/golden/make_golden.py does for the 3-D path:

    python tests/golden2d/make_golden2d.py

Each <case>.npz holds the ini text, the reference's step-0 state, its state after 1 and after N steps (interior cells,
8 variables, float64) and the dt / t values it printed. They pin oracle/mhd2d_oracle.c (tests/test_oracle2d_vs_golden.py);
the CUDA kernels of the 2-D path (next round) will be checked against the same files."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def make_ini2d(n=(16, 12), nstepmax=5, bc=3, riemann="hlld", cfl=0.8, slope_type=2, smallr="1e-8", version=0, tend=10.0):
    bcs = bc if isinstance(bc, (list, tuple)) else [bc] * 4
    bc_txt = "\n".join(f"boundary_type_{nm}={v}" for nm, v in zip(("xmin", "xmax", "ymin", "ymax"), bcs))
    return f"""[run]
solver_name=MHD_Muscl_2D
tEnd={tend}
nStepmax={nstepmax}
nOutput=1
nlog=1
[mesh]
nx={n[0]}
ny={n[1]}
xmin=0.0
xmax=1.0
ymin=0.0
ymax=1.0
{bc_txt}
[hydro]
gamma0=1.666
cfl={cfl}
niter_riemann=10
iorder=2
slope_type={slope_type}
problem=orszag_tang
riemann={riemann}
smallr={smallr}
smallc={smallr}
[output]
outputPrefix=run
outputVtkAscii=false
[other]
implementationVersion={version}
"""


CASES = {
    # name: kwargs of make_ini2d (nstepmax = N)
    "ot2d_16x12": dict(n=(16, 12), nstepmax=6),
    "ot2d_24x24_1e-7": dict(n=(24, 24), nstepmax=8, smallr="1e-7"),  # the floors of settings/test_mhd_orszag_tang_2D.ini
    "ot2d_dirichlet_12x16": dict(n=(12, 16), nstepmax=5, bc=1),
    "ot2d_mixedbc_16x16": dict(n=(16, 16), nstepmax=5, bc=[2, 1, 3, 3]),
    "ot2d_hll_16x12": dict(n=(16, 12), nstepmax=5, riemann="hll"),
    "ot2d_llf_12x12": dict(n=(12, 12), nstepmax=5, riemann="llf"),
    "ot2d_minmod_16x16": dict(n=(16, 16), nstepmax=5, slope_type=1, cfl=0.5),
}


def run(case):
    kw = dict(CASES[case])
    nsteps = kw.pop("nstepmax")
    out = {}
    for tag, ns in (("step1", 1), ("stepN", nsteps)):
        ini = make_ini2d(nstepmax=ns, **kw)
        stdout, states = O.run_reference(ini, threads=4)
        assert len(states) == 2, stdout
        out["init"] = states[0][:, 0]
        out[tag] = states[1][:, 0]
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
        if len(sys.argv) > 1 and case not in sys.argv[1:]:
            continue
        data = run(case)
        np.savez_compressed(os.path.join(here, case + ".npz"), **data)
        print(case, {k: getattr(v, "shape", None) for k, v in data.items()})
