#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

    make -C oracle ref          # builds oracle/_ref/ppkMHD from /root/reference (needs that tree)
    python tests/golden/make_golden.py

For each case the reference (Kokkos-OpenMP build, implementationVersion=0) is run on the ini text
produced by oracle.make_ini() with nStepmax = 1 and nStepmax = N; its binary .vti dumps (interior
cells, 8 conserved variables) are stored as float64 in <case>.npz together with the ini text and the
`dt`/`t` values the reference printed.  Nothing here is computed by this repository's own code:
these files pin both the C oracle (tests/test_oracle_vs_golden.py) and the CUDA path
(tests/test_gpu_parity.py) to the reference bit for bit.
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
LOOP = "[FieldLoop]\nradius=0.3\namplitude=0.001\nvflow=3\ndensity_in=1\n"

CASES = {
    # name: (problem, (nx,ny,nz), nsteps, extra, bounds, cfl, bc[, riemann])
    "ot_16x12x8": ("orszag_tang", (16, 12, 8), 5, OT, None, 0.8, 3),
    "ot2p5d_16x16x4": ("orszag_tang", (16, 16, 4), 5, "", (0, 1, 0, 1, 0, 0.25), 0.8, 3),
    "blast_12x12x12": ("blast", (12, 12, 12), 6, BLAST, None, 0.8, 3),
    "blast_dirichlet_12x10x8": ("blast", (12, 10, 8), 6, BLAST, None, 0.8, 1),
    "blast_neumann_10x12x8": ("blast", (10, 12, 8), 6, BLAST, None, 0.8, 2),
    "blast_mixedbc_12x12x8": ("blast", (12, 12, 8), 6, BLAST, None, 0.8, [1, 2, 3, 3, 2, 1]),
    "fieldloop_24x12x12": ("field_loop", (24, 12, 12), 5, LOOP, (-1, 1, -0.5, 0.5, -0.5, 0.5), 0.4, 3),
    # the other face solvers of riemann_mhd (RiemannSolvers_MHD.h:372-392)
    "blast_hll_12x12x12": ("blast", (12, 12, 12), 6, BLAST, None, 0.8, 3, "hll"),
    "ot_llf_16x12x8": ("orszag_tang", (16, 12, 8), 5, OT, None, 0.8, 3, "llf"),
    "blast_llf_dirichlet_12x10x8": ("blast", (12, 10, 8), 6, BLAST, None, 0.8, 1, "llf"),
    "ot_hll_16x12x8": ("orszag_tang", (16, 12, 8), 5, OT, None, 0.8, 3, "hll"),
    # the other initial conditions of SolverMHDMuscl<3>::init (SolverMHDMuscl.h:653-713)
    "implode_12x12x12": ("implode", (12, 12, 12), 5, "[implode]\nBx_outer=0.3\nBy_inner=0.2\nvx_inner=0.1\n", None, 0.8, 1),
    "kh_robertson_12x12x16": ("kelvin_helmholtz", (12, 12, 16), 5, "[KH]\nd_in=2.0\n", None, 0.8, 3),
    "kh_sine_12x10x16": ("kelvin_helmholtz", (12, 10, 16), 5,
                         "[KH]\nperturbation_sine=true\nperturbation_sine_robertson=false\nd_in=2.0\nw0=0.05\nmode=4\n", None, 0.8, [3, 3, 3, 3, 2, 2]),
    # linear wave on the rotated axis (the reference's convergence test problem): fast wave and Alfven wave
    "wave_fast_16x8x8": ("wave", (16, 8, 8), 5, "[wave]\namplitude=1e-3\ntype=0\n", (0, 3, 0, 1.5, 0, 1.5), 0.8, 3),
    "wave_alfven_16x8x12": ("wave", (16, 8, 12), 5, "[wave]\namplitude=1e-2\ntype=1\n", (0, 3, 0, 1.5, 0, 2.0), 0.8, 3),
}
# Step-0 state only (written to tests/golden/init_only/): the reference's 3-D rotor run turns NaN at the first step
# (By = Bz = 0 exactly: 0/0 in riemann_hlld's star states), so there is nothing to step against.
INIT_ONLY = {
    "rotor_16x16x4": ("rotor", (16, 16, 4), 1, "[rotor]\nr0=0.2\nr1=0.3\n", None, 0.8, 3),
    "rotor_default_24x20x4": ("rotor", (24, 20, 4), 1, "", None, 0.8, 3),
}


def run(case):
    problem, n, nsteps, extra, bounds, cfl, bc = CASES[case][:7]
    riemann = CASES[case][7] if len(CASES[case]) > 7 else "hlld"
    kw = dict(problem=problem, n=n, extra=extra, bounds=bounds, cfl=cfl, bc=bc, nlog=1, tend=10.0, riemann=riemann)
    out = {}
    for tag, ns in (("step1", 1), ("stepN", nsteps)):
        ini = O.make_ini(nstepmax=ns, **kw)
        stdout, states = O.run_reference(ini, threads=4)
        assert len(states) == 2, stdout
        out["init"] = states[0]
        out[tag] = states[1]
        if tag == "stepN":
            log = re.findall(r"time step=\s*(\d+) \(dt=\s*([-0-9.eE+]+) t=\s*([-0-9.eE+]+)\)", stdout)
            out["log_dt"] = np.array([float(m[1]) for m in log])
            out["log_t"] = np.array([float(m[2]) for m in log])
            out["final_time"] = np.array(float(re.search(r"final time is ([-0-9.eE+]+)", stdout).group(1)))
            out["ini"] = np.array(ini)
            out["nsteps"] = np.array(ns)
    return out


def run_init_only(case):
    problem, n, nsteps, extra, bounds, cfl, bc = INIT_ONLY[case][:7]
    ini = O.make_ini(nstepmax=1, problem=problem, n=n, extra=extra, bounds=bounds, cfl=cfl, bc=bc, nlog=1, tend=10.0)
    stdout, states = O.run_reference(ini, threads=4)
    return {"init": states[0], "ini": np.array(ini)}


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    os.makedirs(os.path.join(here, "init_only"), exist_ok=True)
    for case in INIT_ONLY:
        if len(sys.argv) > 1 and case not in sys.argv[1:]:
            continue
        data = run_init_only(case)
        np.savez_compressed(os.path.join(here, "init_only", case + ".npz"), **data)
        print(case, {k: getattr(v, "shape", None) for k, v in data.items()})
    for case in CASES:
        if len(sys.argv) > 1 and case not in sys.argv[1:]:
            continue
        data = run(case)
        np.savez_compressed(os.path.join(here, case + ".npz"), **data)
        print(case, {k: getattr(v, "shape", None) for k, v in data.items()})
