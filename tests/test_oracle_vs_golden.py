"""CPU: the plain-C oracle must reproduce the UNMODIFIED reference bit for bit.

The golden .npz files hold .vti dumps written by oracle/_ref/ppkMHD (tests/golden/make_golden.py).
This is what pins the oracle ("parity pinned"): every later GPU-vs-oracle check inherits it.
"""
import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases, slope_cases


def load(case):
    return np.load(f"{GOLDEN}/{case}.npz")


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_bitwise(case, oracle_mod):
    g = load(case)
    orc = oracle_mod.Oracle(str(g["ini"]))
    assert np.array_equal(orc.interior(), g["init"]), "initial condition differs from the reference"
    dts = [orc.step()]
    assert np.array_equal(orc.interior(), g["step1"]), "state after 1 step differs from the reference"
    orc.run()
    dts += []
    assert orc.iteration == int(g["nsteps"])
    assert np.array_equal(orc.interior(), g["stepN"]), "state after N steps differs from the reference"
    # the reference prints dt with 8 decimals and the final time with 6
    assert abs(dts[0] - g["log_dt"][0]) <= 0.5e-8 + 1e-15
    assert abs(orc.t - float(g["final_time"])) <= 0.5e-6 + 1e-12


@pytest.mark.parametrize("case", slope_cases())
def test_oracle_matches_reference_bitwise_other_limiters(case, oracle_mod):
    """slope_type=1 (minmod) and slope_type=0 (hydro slopes switched off, face-field slopes limited to zero) in 3-D: the branches
    of slope_unsplit_hydro_3d / slope_unsplit_mhd_3d (MHDBaseFunctor3D.h:362-495, 561-668) the main fixtures never take."""
    g = np.load(f"{ROOT}/tests/golden_slope/{case}.npz")
    orc = oracle_mod.Oracle(str(g["ini"]))
    assert np.array_equal(orc.interior(), g["init"]), "initial condition differs from the reference"
    dt1 = orc.step()
    assert np.array_equal(orc.interior(), g["step1"]), "state after 1 step differs from the reference"
    orc.run()
    assert orc.iteration == int(g["nsteps"])
    assert np.array_equal(orc.interior(), g["stepN"]), "state after N steps differs from the reference"
    assert abs(dt1 - g["log_dt"][0]) <= 0.5e-8 + 1e-15
    assert abs(orc.t - float(g["final_time"])) <= 0.5e-6 + 1e-12


def test_known_answers_survey_appendix_d(oracle_mod):
    """SURVEY App. D known answers measured on the reference: OT 32^3 kt=1, 5 steps."""
    O = oracle_mod
    orc = O.Oracle(O.make_ini("orszag_tang", (32, 32, 32), nstepmax=5, extra="[OrszagTang]\nkt=1\n")).run()
    sums, divb = orc.diagnostics()
    assert abs(orc.t - 0.022748) < 1e-6
    assert abs(sums[0] - 7.237524877803631e03) <= 1e-9 * 7.2e3
    assert abs(sums[1] - 1.079356281158038e04) <= 1e-9 * 1.1e4
    assert divb < 5e-14


def test_float_precision_parsing(oracle_mod):
    """ConfigMap::getFloat goes through float."""
    O = oracle_mod
    assert O.parse_float("1.666") == 1.66600000858306884765625
    assert O.parse_float("0.8") == 0.800000011920928955078125
    assert O.parse_float("1e-8") == float(np.float32(1e-8))
    assert O.parse_float("", 0.5) == 0.5
    p = O.params_from_config(O.Config(O.make_ini()))
    assert p.smallp == p.smallc * p.smallc / p.gamma0
    assert (p.isize, p.jsize, p.ksize) == (38, 38, 38)


@pytest.mark.skipif("not __import__('oracle.oracle').oracle.have_reference()")
def test_oracle_vs_live_reference_blast_32(oracle_mod):
    """When the reference binary is present, also compare live (not only through fixtures)."""
    O = oracle_mod
    ini = O.make_ini("blast", (20, 24, 16), nstepmax=4,
                     extra="[blast]\nradius=0.2\npressure_in=10.0\npressure_out=0.1\n")
    _, states = O.run_reference(ini, threads=2)
    orc = O.Oracle(ini).run()
    assert np.array_equal(orc.interior(), states[-1])
