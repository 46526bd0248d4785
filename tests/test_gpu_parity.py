"""GPU parity tests (run with -m gpu on the B200 box). Everything goes through the C ABI of
include/ppkmhd_b200.h; the checker is (a) the golden fixtures written by the unmodified reference and
(b) the plain-C oracle, itself pinned to the reference bit for bit.

Bars (BASELINE.json north_star): exact-arithmetic build => BIT-IDENTICAL conserved variables;
fast (FMA) build => |a-b| <= 1e-12*max(|a|, max_domain|var|) per cell after one step, conserved sums
within 1e-10 relative and max|div B| <= 1e-12 after 100 steps.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases

pytestmark = pytest.mark.gpu

import ppkmhd_b200 as ppk  # noqa: E402

OT = "[OrszagTang]\nkt=1\n"
BLAST = "[blast]\nradius=0.25\ndensity_in=1.0\ndensity_out=1.2\npressure_in=10.0\npressure_out=0.1\n"


def make_solver(ini, exact=True, pipeline=None):
    """pipeline=None keeps the handle's default ("unfused": the schedule that measured fastest at 256^3 and 512^3)."""
    p, t_end, nstep = ppk.params_from_ini(ini, exact=exact)
    s = ppk.Mhd3d(p)
    if pipeline is not None:
        s.set_pipeline(pipeline)
    s.upload(ppk.init_condition_from_ini(ini))
    s.set_time(0.0, t_end, 0)
    return s, nstep


def close_per_cell(a, b, rtol=1e-12):
    """|a-b| <= rtol * max(|a|, max_domain |field|) per cell; the domain maximum is taken over
    the vector a component belongs to (|m|, |B|), since e.g. Bz of the field loop is pure round-off (1e-20)."""
    groups = {0: [0], 1: [1], 2: [2, 3, 4], 3: [2, 3, 4], 4: [2, 3, 4], 5: [5, 6, 7], 6: [5, 6, 7], 7: [5, 6, 7]}
    for v in range(8):
        scale = np.maximum(np.abs(b[v]), max(np.abs(b[w]).max() for w in groups[v]))
        bad = np.abs(a[v] - b[v]) > rtol * scale + 1e-300
        assert not bad.any(), f"var {v}: max abs diff {np.abs(a[v] - b[v]).max():.3e} (field max {np.abs(b[v]).max():.3e})"


@pytest.mark.parametrize("pipeline", ["fused", "fused_split", "unfused", "streamed"])
@pytest.mark.parametrize("case", golden_cases())
def test_exact_mode_bit_identical_to_reference(case, pipeline):
    g = np.load(f"{GOLDEN}/{case}.npz")
    s, nstep = make_solver(str(g["ini"]), exact=True, pipeline=pipeline)
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


@pytest.mark.parametrize("pipeline", ["fused", "unfused", "streamed"])
@pytest.mark.parametrize("case", golden_cases())
def test_fast_mode_within_1e12_of_reference(case, pipeline):
    g = np.load(f"{GOLDEN}/{case}.npz")
    s, nstep = make_solver(str(g["ini"]), exact=False, pipeline=pipeline)
    s.step()
    close_per_cell(s.interior(), g["step1"], 1e-12)
    s.run(nstep - 1)
    close_per_cell(s.interior(), g["stepN"], 1e-11)  # a few steps of round-off growth
    s.close()


@pytest.mark.parametrize("problem,n,extra,bounds,cfl,bc", [
    ("orszag_tang", (48, 40, 36), OT, None, 0.8, 3),
    ("orszag_tang", (64, 36, 40), OT, None, 0.8, 3),  # nx a multiple of 32 + periodic x: the wrapped face / edge column
    ("blast", (40, 40, 40), BLAST, None, 0.8, 3),
    ("field_loop", (64, 32, 32), "[FieldLoop]\nradius=0.3\namplitude=0.001\nvflow=3\ndensity_in=1\n", (-1, 1, -0.5, 0.5, -0.5, 0.5), 0.4, 3),
    # TMA-sized tiles with non-periodic faces: the left-over x column goes through the plain-load kernels (no wrap)
    ("blast", (64, 36, 34), BLAST, None, 0.8, [1, 2, 2, 1, 3, 3]),
    ("implode", (32, 40, 36), "[implode]\nBx_outer=0.3\nBy_inner=0.2\n", None, 0.8, 1),
])
def test_exact_mode_vs_oracle_with_intermediates(problem, n, extra, bounds, cfl, bc, oracle_mod):
    """Larger grids than the fixtures, against the C oracle, including every intermediate array."""
    O = oracle_mod
    ini = O.make_ini(problem, n, nstepmax=4, extra=extra, bounds=bounds, cfl=cfl, tend=10.0, bc=bc)
    orc = O.Oracle(ini)
    s, _ = make_solver(ini, exact=True, pipeline="unfused")   # the pipeline that stores Fluxes_* and Emf
    f, _ = make_solver(ini, exact=True, pipeline="fused")
    m, _ = make_solver(ini, exact=True, pipeline="streamed")
    tl, _ = make_solver(ini, exact=True, pipeline="tiled")   # fused producer + the six Riemann tasks in one launch
    tf, _ = make_solver(ini, exact=False, pipeline="tiled")  # the same in fast arithmetic
    od, _ = make_solver(ini, exact=True, pipeline="ordered")  # unfused producers + the six Riemann tasks in one launch
    for step in range(4):
        orc.step()
        s.step()
        f.step()
        m.step()
        tl.step()
        tf.step()
        od.step()
        assert np.array_equal(od.interior(), orc.interior()), f"ordered pipeline differs at step {step + 1}"
        assert np.array_equal(f.interior(), orc.interior()), f"fused pipeline differs at step {step + 1}"
        assert np.array_equal(m.interior(), orc.interior()), f"streamed pipeline differs at step {step + 1}"
        assert np.array_equal(tl.interior(), orc.interior()), f"tiled pipeline differs at step {step + 1}"
        assert tl.get_time() == s.get_time()
        close_per_cell(tf.interior(), orc.interior(), 1e-12 if step == 0 else 1e-11)
        t, dt, it = s.get_time()
        assert dt == orc.dt and t == orc.t, f"dt/t differ at step {step}: {dt} vs {orc.dt}"
        if step == 0:
            gw = 3
            Q = s.debug_array("Q")
            assert np.array_equal(Q[:, :-1, :-1, :-1], orc.Q[:, :-1, :-1, :-1]), "primitive variables"
            E = s.debug_array("ElecField")
            assert np.array_equal(E[:, 1:-1, 1:-1, 1:-1], orc.scratch_array("ElecField", 3)[:, 1:-1, 1:-1, 1:-1])
            # limited face-field slopes: dA/dy, dA/dz, dB/dx, dB/dz, dC/dx, dC/dy (ComputeMagSlopesFunctor3D)
            dbf = s.debug_array("dbf")
            dA, dB, dC = (orc.scratch_array(nm, 3) for nm in ("DeltaA", "DeltaB", "DeltaC"))
            for comp, want in enumerate((dA[1], dA[2], dB[0], dB[2], dC[0], dC[1])):
                assert np.array_equal(dbf[comp, 1:-1, 1:-1, 1:-1], want[1:-1, 1:-1, 1:-1]), f"face-field slope {comp}"
            # the fused producer writes the same basis and slopes as the three kernels it replaces
            inner = (slice(None), slice(2, -2), slice(2, -2), slice(2, -2))
            assert np.array_equal(tl.debug_array("basis")[inner], s.debug_array("basis")[inner]), "basis of the fused producer"
            assert np.array_equal(tl.debug_array("dbf")[inner], dbf[inner]), "face-field slopes of the fused producer"
            for name in ("Fluxes_x", "Fluxes_y", "Fluxes_z", "Emf"):
                assert np.array_equal(tl.debug_array(name), s.debug_array(name)), name + " of the one-launch Riemann kernel"
                assert np.array_equal(od.debug_array(name), s.debug_array(name)), name + " of the ordered pipeline"
            emf = s.debug_array("Emf")
            eo = orc.scratch_array("Emf", 3)
            nz, ny, nx = n[2], n[1], n[0]
            # each EMF component on the edges the CT update reads
            assert np.array_equal(emf[0, gw:gw + nz, gw:gw + ny + 1, gw:gw + nx + 1], eo[0, gw:gw + nz, gw:gw + ny + 1, gw:gw + nx + 1]), "EMF_z"
            assert np.array_equal(emf[1, gw:gw + nz + 1, gw:gw + ny, gw:gw + nx + 1], eo[1, gw:gw + nz + 1, gw:gw + ny, gw:gw + nx + 1]), "EMF_y"
            assert np.array_equal(emf[2, gw:gw + nz + 1, gw:gw + ny + 1, gw:gw + nx], eo[2, gw:gw + nz + 1, gw:gw + ny + 1, gw:gw + nx]), "EMF_x"
            for d, name in enumerate(("Fluxes_x", "Fluxes_y", "Fluxes_z")):
                F = s.debug_array(name)
                Fo = orc.scratch_array(name, 8)[:5]  # rho, E, normal, t1, t2 momentum fluxes
                sl = [slice(gw, gw + nz + (d == 2)), slice(gw, gw + ny + (d == 1)), slice(gw, gw + nx + (d == 0))]
                assert np.array_equal(F[(slice(None), *sl)], Fo[(slice(None), *sl)]), name
        assert np.array_equal(s.interior(), orc.interior()), f"state differs at step {step + 1}"
    sums, divb = s.diagnostics()
    so, do = orc.diagnostics()
    assert np.allclose(sums, so, rtol=1e-11, atol=1e-9), (sums, so)  # different summation order
    assert abs(divb - do) <= 1e-18 + 1e-12 * do, (divb, do)
    s.close()
    f.close()
    m.close()
    tl.close()
    tf.close()
    od.close()


def test_hundred_steps_conserved_sums_and_divb(oracle_mod):
    """North-star 100-step criterion at 32^3 (oracle run takes ~20 s): sums within 1e-10, div B <= 1e-12."""
    O = oracle_mod
    ini = O.make_ini("orszag_tang", (32, 32, 32), nstepmax=100, extra=OT, tend=10.0)
    orc = O.Oracle(ini).run()
    so, divb_o = orc.diagnostics()
    for exact in (True, False):
        s, nstep = make_solver(ini, exact=exact)
        s.run(nstep)
        t, _, it = s.get_time()
        assert it == 100
        sums, divb = s.diagnostics()
        scale = np.abs(orc.interior()).reshape(8, -1).sum(axis=1)
        assert np.all(np.abs(sums - so) <= 1e-10 * np.maximum(np.abs(so), 1e-10 * scale) + 1e-10 * scale * 1e-2), (sums, so)
        assert divb <= max(1e-12, 4 * divb_o)
        if exact:
            assert t == orc.t and np.array_equal(s.interior(), orc.interior())
        else:
            assert abs(t - orc.t) <= 1e-12
        s.close()


def test_benchmarked_workload_128_vs_oracle(oracle_mod):
    """The workload bench.py times (Orszag-Tang kt=1: a genuinely 3-D flow, every flux and EMF component exercised) at
    128^3 against the oracle: the exact build bit for bit after 1 and 3 steps on the default (unfused), the ordered and the
    tiled pipeline, the fast build (the one that is benchmarked) within 1e-12 per cell after one step."""
    O = oracle_mod
    ini = O.make_ini("orszag_tang", (128, 128, 128), nstepmax=3, extra=OT, tend=10.0)
    orc = O.Oracle(ini)
    orc.step()
    ref1 = orc.interior().copy()
    orc.step()
    orc.step()
    ref3 = orc.interior()
    for pipeline in (None, "ordered", "tiled"):
        s, _ = make_solver(ini, exact=True, pipeline=pipeline)
        s.step()
        assert np.array_equal(s.interior(), ref1), f"exact build, pipeline {pipeline}: step 1 differs from the oracle"
        s.run(2)
        assert s.get_time()[0] == orc.t
        assert np.array_equal(s.interior(), ref3), f"exact build, pipeline {pipeline}: step 3 differs from the oracle"
        s.close()
        f, _ = make_solver(ini, exact=False, pipeline=pipeline)
        f.step()
        close_per_cell(f.interior(), ref1, 1e-12)
        f.run(2)
        close_per_cell(f.interior(), ref3, 1e-11)
        f.close()


def test_benchmarked_workload_256_fast_vs_exact():
    """BASELINE configs[1] exactly as bench.py runs it (Orszag-Tang kt=1, 256^3, fast build, default pipeline) against the
    exact build, which is pinned to the reference bit for bit at every size the oracle can reach: 1e-12 per cell after
    one step (north-star criterion)."""
    from oracle import oracle as O  # ini text helper only

    ini = O.make_ini("orszag_tang", (256, 256, 256), nstepmax=1, extra=OT, tend=10.0)
    e, _ = make_solver(ini, exact=True)
    e.step()
    want = e.interior()
    te, dte, _ = e.get_time()
    e.close()
    for pipeline in (None, "ordered", "tiled"):
        f, _ = make_solver(ini, exact=False, pipeline=pipeline)
        f.step()
        tf, dtf, _ = f.get_time()
        assert abs(dtf - dte) <= 1e-14 * dte
        close_per_cell(f.interior(), want, 1e-12)
        f.close()


def test_hundred_steps_128_conserved_sums_and_divb(oracle_mod):
    """North-star 100-step criterion at the size SURVEY 8(d) names (128^3, kt=1): total mass, momentum, energy and
    magnetic flux within 1e-10 relative of the reference arithmetic (the pinned oracle), max|div B| <= 1e-12; the exact
    build reproduces the oracle's state bit for bit after the 100 steps."""
    O = oracle_mod
    ini = O.make_ini("orszag_tang", (128, 128, 128), nstepmax=100, extra=OT, tend=10.0)
    orc = O.Oracle(ini).run()
    so, divb_o = orc.diagnostics()
    scale = np.abs(orc.interior()).reshape(8, -1).sum(axis=1)
    for exact in (True, False):
        s, nstep = make_solver(ini, exact=exact)
        s.run(nstep)
        t, _, it = s.get_time()
        assert it == 100
        sums, divb = s.diagnostics()
        assert np.all(np.abs(sums - so) <= 1e-10 * np.maximum(np.abs(so), 1e-2 * scale)), (sums, so)
        assert divb <= max(1e-12, 4 * divb_o), (divb, divb_o)
        if exact:
            assert t == orc.t and np.array_equal(s.interior(), orc.interior())
        else:
            assert abs(t - orc.t) <= 1e-12
        s.close()


def test_dropin_executable_matches_reference_vti_bytes(oracle_mod):
    """ppkMHD_b200 <ini> (SolverFactory -> 'MHD_Muscl_3D' -> C ABI) writes the same .vti payload as the
    reference; when the reference binary travelled to this box, compare the files byte for byte."""
    O = oracle_mod
    g = np.load(f"{GOLDEN}/ot_16x12x8.npz")
    ini = str(g["ini"])
    exe = os.path.join(ROOT, "ppkmhd_b200", "bin", "ppkMHD_b200")
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "run.ini"), "w").write(ini)
        out = subprocess.run([exe, "run.ini"], cwd=tmp, capture_output=True, text=True, check=True).stdout
        files = sorted(f for f in os.listdir(tmp) if f.endswith(".vti"))
        assert len(files) == 2, out
        assert np.array_equal(O.read_vti(os.path.join(tmp, files[0])), g["init"])
        assert np.array_equal(O.read_vti(os.path.join(tmp, files[1])), g["stepN"])
        assert "final time is %f" % float(g["final_time"]) in out
        assert "time step=      0 (dt=% 10.8f t=% 10.8f)" % (g["log_dt"][0], 0.0) in out
        if O.have_reference():
            with tempfile.TemporaryDirectory() as tmp2:
                open(os.path.join(tmp2, "run.ini"), "w").write(ini)
                subprocess.run([O.REF_BIN, "run.ini"], cwd=tmp2, capture_output=True, check=True,
                               env=dict(os.environ, OMP_NUM_THREADS="4"))
                for f in files:
                    assert open(os.path.join(tmp, f), "rb").read() == open(os.path.join(tmp2, f), "rb").read(), f


def test_full_size_properties_256():
    """BASELINE configs[1] size (256^3): checks that need no oracle run.
    * 2.5-D Orszag-Tang (kt=0) stays exactly z-invariant in exact mode (every k-plane bit-identical);
    * periodic box: mass, momentum, energy and mean field conserved to round-off; div B at round-off."""
    from oracle import oracle as O  # ini text helper only

    ini = O.make_ini("orszag_tang", (256, 256, 256), nstepmax=3, tend=10.0)
    s, _ = make_solver(ini, exact=True)
    s0, d0 = s.diagnostics()
    s.run(3)
    s1, d1 = s.diagnostics()
    U = s.interior()
    assert np.array_equal(U[:, :1].repeat(256, axis=1), U), "z-invariance broken"
    ncell = 256.0 ** 3
    for v in range(8):
        assert abs(s1[v] - s0[v]) <= 1e-12 * ncell * max(1.0, abs(U[v]).max()), (v, s0[v], s1[v])
    assert d1 <= 1e-11
    # fast build on the same problem stays within round-off of the exact build
    f, _ = make_solver(ini, exact=False)
    f.run(3)
    close_per_cell(f.interior(), U, 1e-11)
    f.close()
    s.close()


def test_full_size_properties_512_blast_and_field_loop():
    """BASELINE configs[2] and [3] at their full size (512^3, 96 GB of device arrays), through size-independent
    properties -- no oracle run at this size (the reference's v0 needs 230 GB for it):
    * blast (strong shocks, HLLD positivity path): density and pressure stay positive, periodic box conserves mass,
      momentum, energy and the mean field to round-off, max|div B| stays at round-off;
    * field-loop advection (div B / EMF accuracy): same conservation, div B at round-off, |B| never exceeds its
      initial maximum by more than round-off growth (no spurious field generation)."""
    import torch

    if torch.cuda.mem_get_info(0)[0] < 110e9:
        pytest.skip("needs ~100 GB of free device memory")
    from oracle import oracle as O  # ini text helper only

    n = 512
    blast = "[blast]\nradius=0.1\ndensity_in=1.0\ndensity_out=1.2\npressure_in=10.0\npressure_out=0.1\n"
    loop = "[FieldLoop]\nradius=0.3\namplitude=0.001\nvflow=3\ndensity_in=1\n"
    for problem, extra, bounds, cfl in (("blast", blast, None, 0.8), ("field_loop", loop, (-1, 1, -0.5, 0.5, -0.5, 0.5), 0.4)):
        ini = O.make_ini(problem, (n, n, n), nstepmax=4, extra=extra, bounds=bounds, cfl=cfl, tend=10.0)
        s, _ = make_solver(ini, exact=False)
        s0, d0 = s.diagnostics()
        s.run(4)
        s1, d1 = s.diagnostics()
        t, dt, it = s.get_time()
        assert it == 4 and dt > 0 and np.isfinite(t)
        scale = [max(abs(s0[v]), 1.0) for v in range(8)]
        for v in range(8):
            # sums of 1.3e8 cells: relative 1e-10 of the sum, or of the sum of magnitudes for the signed ones
            assert abs(s1[v] - s0[v]) <= 1e-10 * max(scale[v], float(n) ** 3 * 1e-3), (problem, v, s0[v], s1[v])
        assert d1 <= 1e-10, (problem, d0, d1)
        if problem == "blast":
            U = s.interior()
            assert U[0].min() > 0.0
            eint = U[1] - 0.5 * (U[2] ** 2 + U[3] ** 2 + U[4] ** 2) / U[0] - 0.5 * (U[5] ** 2 + U[6] ** 2 + U[7] ** 2)
            assert eint.min() > 0.0, "negative internal energy (cell-face field used as cell-centred: a lower bound only)"
            del U, eint
        s.close()


def test_linear_wave_convergence_and_tend_clamp(oracle_mod):
    """The reference's only physics test (test/convergence/: fulltest.sh + plot-results.py): a fast magnetosonic wave on
    the rotated axis of a 3 x 1.5 x 1.5 periodic box runs for one period (tEnd = 0.5) and the L1 distance to the initial
    state must fall at second order. Also the only case that runs to tEnd: the last step's dt is clamped
    (SolverBase.cpp:174-177) and the loop stops on `t >= tEnd - 1e-14` (:199)."""
    O = oracle_mod
    wave = "[wave]\ntype=0\n"

    def ini_for(nx):
        ini = O.make_ini("wave", (nx, nx // 2, nx // 2), nstepmax=10000, extra=wave, bounds=(0, 3, 0, 1.5, 0, 1.5), cfl=0.4, tend=0.5)
        return ini.replace("gamma0=1.666", "gamma0=1.6666666667")

    def run_to_tend(s, t_end):
        while True:
            t, dt, it = s.get_time()
            if t >= t_end - 1e-14:
                return t, it
            s.step()

    def l1(a, b):  # plot-results.py: sqrt(sum over variables of (mean |difference|)^2)
        return float(np.sqrt(sum(np.abs(a[v] - b[v]).mean() ** 2 for v in range(8))))

    # 16 x 8 x 8: the exact build against the oracle, to the end time, bit for bit (t, iteration count and state)
    ini = ini_for(16)
    orc = O.Oracle(ini).run()
    s, _ = make_solver(ini, exact=True)
    t, it = run_to_tend(s, 0.5)
    assert it == orc.iteration and t == orc.t, (it, orc.iteration, t, orc.t)
    assert np.array_equal(s.interior(), orc.interior())
    s.close()
    errs = {}
    for nx in (16, 32, 64):
        s, _ = make_solver(ini_for(nx), exact=False)
        u0 = s.interior().copy()
        run_to_tend(s, 0.5)
        errs[nx] = l1(s.interior(), u0)
        s.close()
    order = np.log2(errs[32] / errs[64])
    assert errs[16] > errs[32] > errs[64] and order > 1.7, (errs, order)


def test_error_paths():
    g = np.load(f"{GOLDEN}/ot_16x12x8.npz")
    ini = str(g["ini"])
    for bad in ("hllc", "approx", "nonsense"):  # the reference leaves the MHD flux unset for these (RiemannSolvers_MHD.h:372-392)
        p, _, _ = ppk.params_from_ini(ini.replace("riemann=hlld", "riemann=" + bad))
        with pytest.raises(ppk.PpkError, match="hlld"):
            ppk.Mhd3d(p)
    p, _, _ = ppk.params_from_ini(ini.replace("implementationVersion=0", "implementationVersion=2"))
    with pytest.raises(ppk.PpkError, match="implementationVersion"):
        ppk.Mhd3d(p)
    # v1 (the reference's atomic-scatter variant of the same arithmetic) is accepted and runs v0's kernels
    s, nstep = make_solver(ini.replace("implementationVersion=0", "implementationVersion=1"), exact=True)
    s.run(nstep)
    assert np.array_equal(s.interior(), g["stepN"])
    s.close()


def test_fast_math_primitives_within_2ulp():
    """The fast build's rcp / sqrt / rsqrt (MUFU seed + one cubic Newton step) against IEEE results."""
    rng = np.random.default_rng(7)
    x = np.concatenate([10.0 ** rng.uniform(-30, 30, 200000), rng.uniform(0.5, 2.0, 100000), [1.0, 2.0, 4.0, 1e-8, 1.66600000858306884765625]])
    rcp, sq, rsq = ppk.selftest_fastmath(x)
    ulp = lambda got, want: np.abs(got - want) / np.spacing(np.abs(want))
    assert ulp(rcp, 1.0 / x).max() <= 2.0, ulp(rcp, 1.0 / x).max()
    assert ulp(sq, np.sqrt(x)).max() <= 2.0, ulp(sq, np.sqrt(x)).max()
    assert ulp(rsq, 1.0 / np.sqrt(x)).max() <= 2.0, ulp(rsq, 1.0 / np.sqrt(x)).max()
    # negative arguments of the reciprocal, and sqrt(0) == 0 (the clamped fast-speed discriminant can vanish)
    rcp, _, _ = ppk.selftest_fastmath(-x[:1000])
    assert ulp(rcp, -1.0 / x[:1000]).max() <= 2.0
    _, sq0, _ = ppk.selftest_fastmath(np.zeros(4))
    assert np.array_equal(sq0, np.zeros(4))


def _slab_worker(rank, world, port, ini, nsteps, exact, out_dir, pipeline=None):
    import sys

    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # CPU side channel for the 128-byte NCCL id only
    import ppkmhd_b200 as P

    p, t_end, _ = P.params_from_ini(ini, rank_z=rank, device=rank, exact=exact)
    s = P.Mhd3d(p)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.tensor(list(P.nccl_unique_id()), dtype=torch.uint8)
    dist.broadcast(idt, 0)
    s.comm_init(bytes(idt.tolist()), world, rank)
    if pipeline is not None:
        s.set_pipeline(pipeline)
    s.upload(P.init_condition_from_ini(ini, rank_z=rank))
    s.set_time(0.0, t_end, 0)
    s.run(nsteps)
    t, dt, it = s.get_time()
    sums, divb = s.diagnostics()
    np.save(os.path.join(out_dir, f"slab{rank}.npy"), s.interior())
    np.save(os.path.join(out_dir, f"meta{rank}.npy"), np.array([t, dt, it, divb] + list(sums)))
    s.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nxy,nzl,pipeline", [(2, (24, 20), 12, None), (4, (24, 20), 12, None), (2, (64, 36), 16, None),
                                                     (2, (64, 36), 16, "ordered"), (4, (512, 512), 64, None)])
def test_z_slabs_bit_identical_to_single_gpu(tmp_path, world, nxy, nzl, pipeline):
    """Decomposition invariance: N z-slabs over NCCL == the undecomposed run, bit for bit, dt included."""
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import oracle as O  # ini text helper only

    # (64, 36): TMA-staged tiles, the wrapped x column and the early halo exchange in the decomposed run;
    # (512, 512) x 64 planes per slab: the production schedule at bench.py's plane size (ordered pipeline, y-slabs for the L2)
    nsteps = 6 if nxy[0] < 512 else 3
    kw = dict(nstepmax=nsteps, extra="[OrszagTang]\nkt=0.5\n", tend=10.0)
    ini_n = O.make_ini("orszag_tang", (*nxy, nzl), mz=world, **kw)
    ini_1 = O.make_ini("orszag_tang", (*nxy, nzl * world), **kw)
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mp.spawn(_slab_worker, args=(world, port, ini_n, nsteps, True, str(tmp_path), pipeline), nprocs=world, join=True)
    s, _ = make_solver(ini_1, exact=True)
    s.run(nsteps)
    t, dt, it = s.get_time()
    sums, divb = s.diagnostics()
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)], axis=1)
    for r in range(world):
        m = np.load(tmp_path / f"meta{r}.npy")
        assert m[0] == t and m[1] == dt and m[2] == it, (r, m[:3], t, dt, it)
    assert np.array_equal(got, s.interior()), "z-slab run differs from the single-GPU run"
    tot = sum(np.load(tmp_path / f"meta{r}.npy")[4:] for r in range(world))
    # (the per-slab sums add up in another order than the single-GPU reduction: round-off of ncell terms of O(1))
    assert np.allclose(tot, sums, rtol=1e-12, atol=1e-15 * got[0].size)
    s.close()


@pytest.mark.parametrize("shape,nblock,bc", [((2, 2, 1), (32, 20, 16), 3), ((1, 2, 2), (64, 12, 10), 3), ((2, 1, 2), (32, 24, 12), [1, 2, 3, 3, 2, 1]),
                                             ((2, 2, 2), (32, 12, 10), 3)])
def test_blocks_bit_identical_to_single_gpu(tmp_path, shape, nblock, bc):
    """Pencil / block decomposition ([mpi] mx, my > 1): x and y faces through the packed exchange
    (k_face_copy + grouped ncclSend / ncclRecv), z planes as before, X -> Y -> Z like make_boundaries_mpi
    (SolverBase.cpp:610-693). The decomposed run must equal the undecomposed one bit for bit, dt included, with periodic
    and with mixed physical boundaries (edges and corners are where a wrong exchange order shows)."""
    import socket

    import torch
    import torch.multiprocessing as mp

    mx, my, mz = shape
    world = mx * my * mz
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import oracle as O  # ini text helper only

    nsteps = 5
    kw = dict(nstepmax=nsteps, extra="[OrszagTang]\nkt=0.5\n", tend=10.0, bc=bc)
    ini_n = O.make_ini("orszag_tang", nblock, mx=mx, my=my, mz=mz, **kw)
    ini_1 = O.make_ini("orszag_tang", (nblock[0] * mx, nblock[1] * my, nblock[2] * mz), **kw)
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mp.spawn(_slab_worker, args=(world, port, ini_n, nsteps, True, str(tmp_path), None), nprocs=world, join=True)
    s, _ = make_solver(ini_1, exact=True)
    s.run(nsteps)
    t, dt, it = s.get_time()
    want = s.interior()
    s.close()
    got = np.empty_like(want)
    nx, ny, nz = nblock
    for r in range(world):
        cz, cy, cx = r % mz, (r // mz) % my, r // (mz * my)
        got[:, cz * nz:(cz + 1) * nz, cy * ny:(cy + 1) * ny, cx * nx:(cx + 1) * nx] = np.load(tmp_path / f"slab{r}.npy")
        m = np.load(tmp_path / f"meta{r}.npy")
        assert m[0] == t and m[1] == dt and m[2] == it, (r, m[:3], t, dt, it)
    assert np.array_equal(got, want), "block-decomposed run differs from the single-GPU run"


def test_split_phase_transfers_and_default_schedule():
    """ppk_mhd3d_stage_upload / _stage_swap / _stage_download (pipelined host transfers on one handle, what bench.py's e2e
    leg drives) return exactly what upload -> step -> download returns, batch after batch, while four device arrays
    rotate; the default schedule is the measured one (unfused) at every plane size and get_pipeline reports what was set."""
    import torch

    from oracle import oracle as O  # ini text helper only

    ini = O.make_ini("orszag_tang", (64, 36, 20), nstepmax=1, extra=OT, tend=10.0)
    ref, _ = make_solver(ini, exact=True)
    assert ref.pipeline() == "unfused"
    u0 = ppk.init_condition_from_ini(ini)
    ref.step()
    want1 = ref.download().copy()       # one step from the initial state
    ref.step()
    want2 = ref.download().copy()       # two steps (same handle: U / U2 parity flips)
    ref.close()

    p, t_end, _ = ppk.params_from_ini(ini, exact=True)
    s = ppk.Mhd3d(p)
    s.set_time(0.0, t_end, 0)
    pin = [torch.from_numpy(u0.copy()).pin_memory(), torch.from_numpy(want1.copy()).pin_memory()]
    out = [torch.empty(p.shape, dtype=torch.float64).pin_memory() for _ in range(6)]
    # batches alternate between the initial state and the one-step state: results must be want1 / want2 alternately
    s.stage_upload(pin[0].data_ptr())
    for b in range(6):
        s.stage_swap()
        s.step()
        s.stage_download(out[b].data_ptr())
        if b + 1 < 6:
            s.stage_upload(pin[(b + 1) % 2].data_ptr())
    s.synchronize()
    gw = 3
    inner = (slice(None), slice(gw, -gw), slice(gw, -gw), slice(gw, -gw))
    for b in range(6):
        want = want1 if b % 2 == 0 else want2
        assert np.array_equal(out[b].numpy()[inner], want[inner]), f"batch {b} differs"
    s.close()

    big = O.make_ini("orszag_tang", (384, 384, 8), nstepmax=1, extra=OT, tend=10.0)
    pb, _, _ = ppk.params_from_ini(big, exact=False)
    sb = ppk.Mhd3d(pb)
    assert sb.pipeline() == "unfused"
    sb.set_pipeline("ordered")
    assert sb.pipeline() == "ordered"
    sb.close()
