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


def make_ini2d(n=(16, 12), nstepmax=5, bc=3, riemann="hlld", cfl=0.8, slope_type=2, smallr="1e-8", version=0, tend=10.0,
               problem="orszag_tang", extra="", bounds=(0.0, 1.0, 0.0, 1.0), gamma0="1.666"):
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
xmin={bounds[0]}
xmax={bounds[1]}
ymin={bounds[2]}
ymax={bounds[3]}
{bc_txt}
[hydro]
gamma0={gamma0}
cfl={cfl}
niter_riemann=10
iorder=2
slope_type={slope_type}
problem={problem}
riemann={riemann}
smallr={smallr}
smallc={smallr}
[output]
outputPrefix=run
outputVtkAscii=false
[other]
implementationVersion={version}
{extra}"""


CASES = {
    # name: kwargs of make_ini2d (nstepmax = N)
    "ot2d_16x12": dict(n=(16, 12), nstepmax=6),
    "ot2d_24x24_1e-7": dict(n=(24, 24), nstepmax=8, smallr="1e-7"),  # the floors of settings/test_mhd_orszag_tang_2D.ini
    "ot2d_dirichlet_12x16": dict(n=(12, 16), nstepmax=5, bc=1),
    "ot2d_mixedbc_16x16": dict(n=(16, 16), nstepmax=5, bc=[2, 1, 3, 3]),
    "ot2d_hll_16x12": dict(n=(16, 12), nstepmax=5, riemann="hll"),
    "ot2d_llf_12x12": dict(n=(12, 12), nstepmax=5, riemann="llf"),
    "ot2d_minmod_16x16": dict(n=(16, 16), nstepmax=5, slope_type=1, cfl=0.5),
    # the other 2-D problems with a shipped .ini (test_mhd_blast_2D, test_mhd_rotor, mhd_fieldloop2d, ..._kelvin_helmholtz_2D_mhd)
    "blast2d_16x16": dict(n=(16, 16), nstepmax=6, problem="blast",
                          extra="[blast]\nradius=0.25\ndensity_in=1.0\ndensity_out=1.2\npressure_in=10.0\npressure_out=0.1\n"),
    "blast2d_neumann_12x16": dict(n=(12, 16), nstepmax=6, problem="blast", bc=2,
                                  extra="[blast]\nradius=0.25\ndensity_in=1.0\ndensity_out=1.2\npressure_in=10.0\npressure_out=0.1\n"),
    # settings/test_mhd_rotor.ini at 32^2 (with gamma0 = 1.666, u0 = 2 and periodic faces the reference itself turns NaN)
    "rotor2d_32x32": dict(n=(32, 32), nstepmax=8, problem="rotor", bc=2, cfl=0.4, gamma0="1.4",
                          extra="[rotor]\nr0=0.1\nr1=0.115\nu0=1.0\np0=1.5\n"),
    "fieldloop2d_24x12": dict(n=(24, 12), nstepmax=6, problem="field_loop", cfl=0.4, bounds=(-1.0, 1.0, -0.5, 0.5),
                              extra="[FieldLoop]\nradius=0.3\namplitude=0.001\nvflow=3\ndensity_in=1\n"),
    "implode2d_16x16": dict(n=(16, 16), nstepmax=6, problem="implode", bc=1, extra="[implode]\nBx_outer=0.3\nBy_inner=0.2\nvx_inner=0.1\n"),
    "kh2d_robertson_16x16": dict(n=(16, 16), nstepmax=5, problem="kelvin_helmholtz", extra="[KH]\nd_in=2.0\n"),
    "kh2d_sine_16x12": dict(n=(16, 12), nstepmax=5, problem="kelvin_helmholtz", bc=[3, 3, 2, 2],
                            extra="[KH]\nperturbation_sine=true\nperturbation_sine_robertson=false\nd_in=2.0\nw0=0.05\nmode=4\n"),
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
