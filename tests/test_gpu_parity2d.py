"""GPU parity tests of the 2-D path (MHD_Muscl_2D, implementationVersion 0) through the C ABI (ppk_mhd2d_*): the exact
build must reproduce the fixtures written by the unmodified reference bit for bit, the fast build within 1e-12 per cell
after one step (tests/golden2d/make_golden2d.py; the same files pin oracle/mhd2d_oracle.c on the CPU)."""
import os

import numpy as np
import pytest

from conftest import ROOT

import ppkmhd_b200 as ppk

pytestmark = pytest.mark.gpu
GOLDEN2D = os.path.join(ROOT, "tests", "golden2d")


def cases2d():
    return sorted(f[:-4] for f in os.listdir(GOLDEN2D) if f.endswith(".npz"))


def make_solver2d(ini, exact=True):
    p, t_end, nstep = ppk.params_from_ini(ini, device=0, exact=exact)
    s = ppk.Mhd2d(p)
    s.upload(ppk.init_condition_2d_from_ini(ini))
    s.set_time(0.0, t_end, 0)
    return s, nstep


@pytest.mark.parametrize("case", cases2d())
def test_exact_mode_2d_bit_identical_to_reference(case):
    g = np.load(f"{GOLDEN2D}/{case}.npz")
    s, nstep = make_solver2d(str(g["ini"]), exact=True)
    assert np.array_equal(s.interior(), g["init"])
    s.step()
    t, dt, it = s.get_time()
    assert it == 1 and abs(dt - g["log_dt"][0]) <= 0.5e-8 + 1e-15
    assert np.array_equal(s.interior(), g["step1"]), "step 1 differs from the reference"
    s.run(nstep - 1)
    t, dt, it = s.get_time()
    assert it == nstep and abs(t - float(g["final_time"])) <= 0.5e-6 + 1e-12
    assert np.array_equal(s.interior(), g["stepN"]), f"step {nstep} differs from the reference"
    s.close()


@pytest.mark.parametrize("case", cases2d())
def test_fast_mode_2d_within_1e12_of_reference(case):
    g = np.load(f"{GOLDEN2D}/{case}.npz")
    s, nstep = make_solver2d(str(g["ini"]), exact=False)
    s.step()
    a, b = s.interior(), g["step1"]
    for v in range(8):
        scale = max(np.abs(b[v]).max(), np.abs(b[[2, 3, 4]]).max() if v in (2, 3, 4) else 0.0, np.abs(b[[5, 6, 7]]).max() if v in (5, 6, 7) else 0.0)
        assert np.abs(a[v] - b[v]).max() <= 1e-12 * max(scale, 1e-300), (v, np.abs(a[v] - b[v]).max(), scale)
    s.close()


def test_config0_orszag_tang_256_against_oracle(oracle_mod):
    """BASELINE configs[0] (settings/test_mhd_orszag_tang_2D.ini: Orszag-Tang 256^2, periodic, floors 1e-7) with the v0
    formulation: 10 steps of the exact build against the 2-D oracle, bit for bit."""
    import sys

    sys.path.insert(0, GOLDEN2D)
    from make_golden2d import make_ini2d

    ini = make_ini2d(n=(256, 256), nstepmax=10, smallr="1e-7")
    orc = oracle_mod.Oracle2D(ini).run()
    s, nstep = make_solver2d(ini, exact=True)
    s.run(nstep)
    t, dt, it = s.get_time()
    assert it == orc.iteration and t == orc.t
    assert np.array_equal(s.interior(), orc.interior())
    s.close()


def test_executable_2d_writes_the_reference_vti(oracle_mod):
    """`ppkMHD_b200 run.ini` with solver_name=MHD_Muscl_2D (SolverFactory -> SolverMHDMusclCuda2D -> ppk_mhd2d_*): the .vti
    files hold the reference's states and, where the reference binary is present, are byte-identical to its files."""
    import subprocess
    import tempfile

    O = oracle_mod
    g = np.load(f"{GOLDEN2D}/ot2d_16x12.npz")
    ini = str(g["ini"])
    exe = os.path.join(ROOT, "ppkmhd_b200", "bin", "ppkMHD_b200")
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "run.ini"), "w").write(ini)
        out = subprocess.run([exe, "run.ini"], cwd=tmp, capture_output=True, text=True, check=True).stdout
        files = sorted(f for f in os.listdir(tmp) if f.endswith(".vti"))
        assert len(files) == 2, out
        assert np.array_equal(O.read_vti(os.path.join(tmp, files[0]))[:, 0], g["init"])
        assert np.array_equal(O.read_vti(os.path.join(tmp, files[1]))[:, 0], g["stepN"])
        assert "final time is %f" % float(g["final_time"]) in out
        assert "time step=      0 (dt=% 10.8f t=% 10.8f)" % (g["log_dt"][0], 0.0) in out
        if O.have_reference():
            with tempfile.TemporaryDirectory() as tmp2:
                open(os.path.join(tmp2, "run.ini"), "w").write(ini)
                subprocess.run([O.REF_BIN, "run.ini"], cwd=tmp2, capture_output=True, check=True, env=dict(os.environ, OMP_NUM_THREADS="4"))
                for f in files:
                    assert open(os.path.join(tmp, f), "rb").read() == open(os.path.join(tmp2, f), "rb").read(), f


@pytest.mark.parametrize("exact", [True, False])
def test_config0_v2_fixture_within_1e12(exact):
    """BASELINE configs[0] as shipped (implementationVersion=2): the GPU path runs the v0 formulation and must agree with
    the reference's v2 output within 1e-12 of the field maximum (tests/golden2d_v2/README.md)."""
    g = np.load(os.path.join(ROOT, "tests", "golden2d_v2", "ot2d_v2_64x64.npz"))
    s, nstep = make_solver2d(str(g["ini"]), exact=exact)   # the ini says implementationVersion=2
    assert np.array_equal(s.interior(), g["init"])
    s.run(nstep)
    t, dt, it = s.get_time()
    assert it == int(g["nsteps"]) and abs(t - float(g["final_time"])) <= 0.5e-6 + 1e-12
    assert np.abs(s.interior() - g["stepN"]).max() <= 1e-12 * np.abs(g["stepN"]).max()
    s.close()
