"""HDF5 output + restart (src/utils/io/IO_HDF5.h: Save_HDF5 :73-526, Load_HDF5 :1537-2153; SolverMHDMuscl.h:615-643) through the
run-time binding of ppkmhd_b200/host/IO_HDF5.cpp. This image has no libhdf5, so the library the binding finds here is
tests/mock_hdf5/mock_hdf5.c: the same ~30 C entry points with the semantics the writer / reader rely on (files, datasets
written / read through a memory-space hyperslab, scalar and string attributes) over a trivial container. What is pinned: the
call sequence, dataset and attribute names, dimensions and the start / count of the ghost-stripping selections, the restart
flow. What is NOT pinned: the HDF5 file format itself (that is libhdf5's job)."""
import os
import struct
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

MOCK_SRC = os.path.join(ROOT, "tests", "mock_hdf5", "mock_hdf5.c")


@pytest.fixture(scope="module")
def mock_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mockhdf5") / "libmockhdf5.so")
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-O1", "-w", "-o", out, MOCK_SRC])
    return out


def read_mock_dataset(path):
    blob = open(path, "rb").read()
    rank = struct.unpack_from("<i", blob, 0)[0]
    dims = struct.unpack_from("<3Q", blob, 4)[:rank]
    return np.frombuffer(blob, dtype="<f8", offset=4 + 24).reshape(dims)


def run_py(code, env_extra, cwd):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=cwd, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("ghosts", [False, True])
def test_hdf5_write_and_restart_read_round_trip(tmp_path, mock_lib, ghosts):
    """Host code only (no GPU): SolverBase::save_data with [output] hdf5_enabled writes one dataset per variable with the ghost
    zones stripped (or kept: ghostIncluded) plus the reference's attributes; IO_ReadWrite::load_data reads it back."""
    code = f"""
        import sys; sys.path.insert(0, {ROOT!r})
        import numpy as np
        import ppkmhd_b200 as ppk
        from oracle import oracle as O
        assert ppk.hdf5_available()
        nx, ny, nz = 10, 6, 4
        ini = O.make_ini("orszag_tang", (nx, ny, nz)).replace("outputPrefix=run", "outputDir={tmp_path}\\noutputPrefix=rt\\nhdf5_enabled=true\\nghostIncluded={'true' if ghosts else 'false'}")
        rng = np.random.default_rng(5)
        U = rng.standard_normal((8, nz + 6, ny + 6, nx + 6))
        np.save("{tmp_path}/U.npy", U)
        ppk.save_data_from_ini(ini, U, 3)
        back = np.full_like(U, -7.0)
        ini_r = ini.replace("[mesh]", "restart_enabled=true\\nrestart_filename={tmp_path}/rt_0000003.h5\\n[mesh]")
        step, t = ppk.load_data_from_ini(ini_r, back)
        np.save("{tmp_path}/back.npy", back)
        print("STEP", step, t)
    """
    out = run_py(code, {"PPK_HDF5_LIB": mock_lib}, str(tmp_path))
    assert "STEP 3 0.0" in out
    U, back = np.load(tmp_path / "U.npy"), np.load(tmp_path / "back.npy")
    h5 = tmp_path / "rt_0000003.h5"
    names = ["rho", "energy", "rho_vx", "rho_vy", "rho_vz", "bx", "by", "bz"]
    inner = (slice(3, -3),) * 3
    for v, nm in enumerate(names):
        d = read_mock_dataset(h5 / f"{nm}.dset")
        want = U[v] if ghosts else U[v][inner]
        assert d.shape == want.shape and np.array_equal(d, want), nm      # slowest dimension first: (nz, ny, nx)
    for attr, fmt, want in (("time step", "<i", 3), ("nx", "<i", 10), ("ny", "<i", 6), ("nz", "<i", 4), ("ghost zone included", "<i", int(ghosts)),
                            ("total time", "<d", 0.0)):
        assert struct.unpack(fmt, open(h5 / f"{attr}.attr", "rb").read())[0] == want, attr
    assert len(open(h5 / "creation date.attr").read()) >= 10
    if ghosts:
        assert np.array_equal(back, U)
    else:  # only the interior is in the file: the ghost layers keep what the caller had
        assert np.array_equal(back[(slice(None),) + inner], U[(slice(None),) + inner])
        mask = np.ones(U.shape, bool)
        mask[(slice(None),) + inner] = False
        assert np.all(back[mask] == -7.0)
    # a file of another resolution is refused
    code2 = f"""
        import sys; sys.path.insert(0, {ROOT!r})
        import numpy as np
        import ppkmhd_b200 as ppk
        from oracle import oracle as O
        ini = O.make_ini("orszag_tang", (12, 6, 4)).replace("[mesh]", "restart_enabled=true\\nrestart_filename={tmp_path}/rt_0000003.h5\\n[mesh]").replace("outputPrefix=run", "hdf5_enabled=true")
        try:
            ppk.load_data_from_ini(ini, np.zeros((8, 10, 12, 18)))
            print("LOADED")
        except ppk.PpkError as e:
            print("REFUSED", e)
    """
    assert "REFUSED" in run_py(code2, {"PPK_HDF5_LIB": mock_lib}, str(tmp_path))


@pytest.mark.gpu
def test_restart_run_continues_bit_identically(tmp_path, mock_lib):
    """SolverMHDMuscl<dim>::init_restart (SolverMHDMuscl.h:615-643) through the ppkMHD_b200 executable: 3 steps, the HDF5 output
    of the last one, a restarted run of 3 more steps == an uninterrupted 6-step run, bit for bit (exact build)."""
    from oracle import oracle as O

    exe = os.path.join(ROOT, "ppkmhd_b200", "bin", "ppkMHD_b200")
    env = dict(os.environ, PPK_HDF5_LIB=mock_lib)
    base = str(np.load(f"{GOLDEN}/ot_16x12x8.npz")["ini"])

    def ini_for(nsteps, prefix, extra_run=""):
        t = base.replace("nStepmax=", "nStepmaxOld=").replace("[run]", f"[run]\nnStepmax={nsteps}\n{extra_run}")
        t = t.replace("outputPrefix=", "outputPrefixOld=").replace("[output]", f"[output]\noutputPrefix={prefix}\nhdf5_enabled=true")
        return t

    def run(ini, name):
        open(tmp_path / name, "w").write(ini)
        r = subprocess.run([exe, name], cwd=tmp_path, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        return r.stdout

    run(ini_for(6, "whole"), "whole.ini")
    run(ini_for(3, "first"), "first.ini")
    h5 = sorted(f for f in os.listdir(tmp_path) if f.startswith("first_") and f.endswith(".h5"))
    out = run(ini_for(3, "second", f"restart_enabled=true\nrestart_filename={tmp_path}/{h5[-1]}"), "second.ini")
    assert "This is a restarted run" in out
    whole = sorted(f for f in os.listdir(tmp_path) if f.startswith("whole_") and f.endswith(".vti"))
    second = sorted(f for f in os.listdir(tmp_path) if f.startswith("second_") and f.endswith(".vti"))
    a, b = O.read_vti(str(tmp_path / whole[-1])), O.read_vti(str(tmp_path / second[-1]))
    assert np.array_equal(a, b), "restarted run differs from the uninterrupted one"
    assert (tmp_path / "whole.xmf").exists()  # the Xdmf wrapper of the series (main.cpp:163-170)
