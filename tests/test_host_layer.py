"""CPU: the C++ host layer (ConfigMap / HydroParams / initial conditions) and the C-ABI surface.
No compute call is made here (there is no GPU on the CPU box)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases, init_only_cases

import ppkmhd_b200 as ppk
from ppkmhd_b200 import capi


def test_library_exports_every_declared_symbol():
    """Every function declared in include/*.h is exported by the shared library."""
    declared = set()
    for hdr in ("ppkmhd_b200.h", "ppkmhd_b200_host.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(ppk_[a-z0-9_]+)\s*\(", text))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    L = C.CDLL(ppk.lib_path())
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} not exported"
    assert b"sm_100a" in ppk.load_library().ppk_version_string()


@pytest.mark.parametrize("case", golden_cases())
def test_params_match_oracle_and_float_parsing(case, oracle_mod):
    ini = str(np.load(f"{GOLDEN}/{case}.npz")["ini"])
    p, t_end, nstep = ppk.params_from_ini(ini)
    o = oracle_mod.params_from_config(oracle_mod.Config(ini))
    for f in ("nx", "ny", "nz", "dx", "dy", "dz", "xmin", "xmax", "ymin", "ymax", "zmin", "zmax",
              "gamma0", "cfl", "slope_type", "smallr", "smallc", "smallp"):
        assert getattr(p, f) == getattr(o, f), f
    assert list(p.boundary_type) == list(o.bc)
    assert p.gamma0 == 1.66600000858306884765625  # float-precision parse
    assert p.ghost_width == 3 and p.riemann_solver == o.riemann and p.implementation_version == 0
    assert p.riemann_solver == {"hll": 2, "llf": 1}.get(case.split("_")[1], 4)
    assert t_end == 10.0 and nstep == int(np.load(f"{GOLDEN}/{case}.npz")["nsteps"])


@pytest.mark.parametrize("case", golden_cases() + init_only_cases())
def test_initial_condition_bitwise(case, oracle_mod):
    g = np.load(f"{GOLDEN}/{case}.npz")
    ini = str(g["ini"])
    U = ppk.init_condition_from_ini(ini)
    assert np.array_equal(U[:, 3:-3, 3:-3, 3:-3], g["init"]), "differs from the reference's step-0 .vti"
    orc = oracle_mod.Oracle(ini)
    Uo = oracle_mod.init_problem(orc.p, orc.cfg)
    assert np.array_equal(U, Uo), "ghost cells differ from the oracle's initial array"


def test_slab_params(oracle_mod):
    """[mpi] mz slabs: dz uses the global extent nz*mz (HydroParams.cpp:400-402); cell centres shift."""
    ini = oracle_mod.make_ini("orszag_tang", (16, 12, 8), extra="[OrszagTang]\nkt=1\n", mz=2)
    p0, _, _ = ppk.params_from_ini(ini, rank_z=0)
    p1, _, _ = ppk.params_from_ini(ini, rank_z=1)
    assert p0.mz == 2 and p1.rank_z == 1 and p0.dz == 1.0 / 16
    U1 = ppk.init_condition_from_ini(ini, rank_z=1)
    orc1 = oracle_mod.Oracle(ini, rank_pos=(0, 0, 1))
    assert np.array_equal(U1, oracle_mod.init_problem(orc1.p, orc1.cfg))
    # the two slabs are the two halves of the undecomposed 16x12x16 problem
    whole = ppk.init_condition_from_ini(oracle_mod.make_ini("orszag_tang", (16, 12, 16), extra="[OrszagTang]\nkt=1\n"))
    assert np.array_equal(U1[:, 3:-3, 3:-3, 3:-3], whole[:, 3 + 8:-3, 3:-3, 3:-3])


def test_configmap_quirks():
    ini = "[Run]\nTEND = 0.5 ; comment\nnstepmax=0x10\n[mesh]\nnx=8\nny=8\nnz=8\n# c\n[hydro]\nriemann=hlld\nproblem=FieldLoop\n"
    p, t_end, nstep = ppk.params_from_ini("[run]\nsolver_name=MHD_Muscl_3D\n" + ini)
    assert t_end == 0.5 and nstep == 16 and p.nx == 8
    assert list(p.boundary_type) == [1] * 6  # default Dirichlet (HydroParams.cpp:145-156)


def test_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p, _, _ = ppk.params_from_ini(str(np.load(f"{GOLDEN}/ot_16x12x8.npz")["ini"]))
    with pytest.raises(ppk.PpkError):
        ppk.Mhd3d(p)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the unmodified reference on the host cores) needs no GPU: one JSON line on stdout with the
    contract's keys, nothing else."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-n", "24"], capture_output=True, text=True, timeout=300, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "Mcell-updates/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_initial_condition_2d_bitwise(oracle_mod):
    """Host-layer 2-D initial conditions (Orszag-Tang, blast, rotor, field loop, Kelvin-Helmholtz) == the reference's step-0
    .vti == the 2-D oracle, ghosts included."""
    g2 = os.path.join(ROOT, "tests", "golden2d")
    for f in sorted(os.listdir(g2)):
        if not f.endswith(".npz"):
            continue
        g = np.load(os.path.join(g2, f))
        ini = str(g["ini"])
        U = ppk.init_condition_2d_from_ini(ini)
        assert np.array_equal(U[:, 3:-3, 3:-3], g["init"]), f
        orc = oracle_mod.Oracle2D(ini)
        Uo = np.zeros_like(U)
        oracle_mod.init_problem_2d(orc.p, orc.cfg, Uo)
        assert np.array_equal(U, Uo), f


def test_pvti_pieces_reassemble_to_the_single_vti(tmp_path):
    """Output contract of a decomposed run (src/utils/io/IO_VTK.cpp:630-853 pieces, :860-1008 .pvti header): three z-slabs
    written through the host layer's save_data must (a) carry the header text the reference writes, line for line, (b) name
    their pieces with the extents of the MPI coordinates, and (c) reassemble, piece by piece, to exactly the payload of the
    single .vti an undecomposed run writes for the same global state. Host code only: no GPU."""
    import re

    import ppkmhd_b200 as ppk
    from oracle import oracle as O

    nx, ny, nzl, mz, gw = 10, 6, 4, 3, 3
    rng = np.random.default_rng(11)
    glob = rng.standard_normal((8, nzl * mz + 2 * gw, ny + 2 * gw, nx + 2 * gw))
    ini_1 = O.make_ini("orszag_tang", (nx, ny, nzl * mz), bounds=(0.0, 1.0, 0.0, 0.5, -1.0, 2.0)).replace("outputPrefix=run", f"outputDir={tmp_path}\noutputPrefix=one")
    ini_n = O.make_ini("orszag_tang", (nx, ny, nzl), mz=mz, bounds=(0.0, 1.0, 0.0, 0.5, -1.0, 2.0)).replace("outputPrefix=run", f"outputDir={tmp_path}\noutputPrefix=slab")
    ppk.save_data_from_ini(ini_1, glob, 7)
    for r in range(mz):
        ppk.save_data_from_ini(ini_n, glob[:, r * nzl:r * nzl + nzl + 2 * gw], 7, rank_z=r)
    whole = O.read_vti(str(tmp_path / "one_0000007.vti"))
    assert np.array_equal(whole, glob[:, gw:-gw, gw:-gw, gw:-gw])

    # (a) header, line by line (write_pvti_header, IO_VTK.cpp:905-1003); dx = 1/10, dy = 0.5/6, dz = 3/12 as operator<< prints them
    p, _, _ = ppk.params_from_ini(ini_n)
    fmt = lambda x: np.format_float_positional(x, precision=6, unique=True, fractional=False, trim="-")
    names = ["rho", "energy", "rho_vx", "rho_vy", "rho_vz", "bx", "by", "bz"]
    want = ['<?xml version="1.0"?>',
            '<VTKFile type="PImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
            f'  <PImageData WholeExtent="0 {nx} 0 {ny} 0 {mz * nzl}" GhostLevel="0" Origin="0 0 -1" Spacing="{fmt(p.dx)} {fmt(p.dy)} {fmt(p.dz)}">',
            '    <PCellData Scalars="Scalars_">'] + [f'      <PDataArray type="Float64" Name="{n}"/>' for n in names] + ['    </PCellData>']
    want += [f' <Piece Extent="0 {nx} 0 {ny} {r * nzl} {r * nzl + nzl} " Source="slab_time0000007_mpi{r:05d}.vti"/>' for r in range(mz)]
    want += ['</PImageData>', '</VTKFile>']
    got = open(tmp_path / "slab_time0000007.pvti").read().splitlines()
    assert got == want, "\n".join(got)

    # (b) + (c): follow the .pvti to its pieces and rebuild the global payload
    rebuilt = np.full_like(whole, np.nan)
    for ln in got:
        m = re.match(r' <Piece Extent="(\d+) (\d+) (\d+) (\d+) (\d+) (\d+) " Source="([^"]+)"/>', ln)
        if not m:
            continue
        x0, x1, y0, y1, z0, z1 = (int(m.group(i)) for i in range(1, 7))
        blob = open(tmp_path / m.group(7), "rb").read()
        head = blob[:blob.index(b"<AppendedData")].decode()
        assert '<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64">' in head
        assert f'WholeExtent="{x0} {x1} {y0} {y1} {z0} {z1}"' in head and f'<Piece Extent="{x0} {x1} {y0} {y1} {z0} {z1}">' in head
        pos = blob.index(b"_", blob.index(b"<AppendedData")) + 1
        for v in range(8):
            nbytes = int(np.frombuffer(blob, dtype="<u8", count=1, offset=pos)[0])
            assert nbytes == (x1 - x0) * (y1 - y0) * (z1 - z0) * 8
            rebuilt[v, z0:z1, y0:y1, x0:x1] = np.frombuffer(blob, dtype="<f8", count=nbytes // 8, offset=pos + 8).reshape(z1 - z0, y1 - y0, x1 - x0)
            pos += 8 + nbytes
    assert np.array_equal(rebuilt, whole)


def test_xdmf_wrapper_text_and_hdf5_error_path(tmp_path, monkeypatch):
    """writeXdmfForHdf5Wrapper (src/utils/io/IO_HDF5.cpp:16-378) is plain text: the file must read, line for line, like the
    reference's (header :138-143, per-output grid :176-186, geometry :208-227, one attribute block per variable in the
    order rho, energy, rho_vx, rho_vy, rho_vz, bx, by, bz :229-360, footer :363-370). libhdf5 itself is looked up at run time:
    where it is absent (this image), an HDF5 output request is reported as PPK_ERR_UNSUPPORTED, never silently dropped."""
    import ppkmhd_b200 as ppk
    from oracle import oracle as O

    monkeypatch.chdir(tmp_path)
    nx, ny, nz = 10, 6, 4
    ini = O.make_ini("orszag_tang", (nx, ny, nz)).replace("outputPrefix=run", f"outputDir={tmp_path}\noutputPrefix=ot3d\nhdf5_enabled=true")
    ppk.write_xdmf_from_ini(ini, 2)
    got = open(tmp_path / "ot3d.xmf").read().splitlines()
    names = ["rho", "energy", "rho_vx", "rho_vy", "rho_vz", "bx", "by", "bz"]
    want = ['<?xml version="1.0" ?>', '<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>',
            '<Xdmf xmlns:xi="http://www.w3.org/2003/XInclude" Version="2.2">', '  <Domain>',
            '    <Grid Name="TimeSeries" GridType="Collection" CollectionType="Temporal">']
    for step in range(3):  # outputs 0 .. totalNumberOfSteps inclusive (:160-163)
        base = f"ot3d_{step:07d}"
        want += [f'    <Grid Name="{base}" GridType="Uniform">', f'    <Time Value="{step}" />',
                 f'      <Topology TopologyType="3DCoRectMesh" NumberOfElements="{nz} {ny} {nx}"/>',
                 '    <Geometry Type="ORIGIN_DXDYDZ">']
        for what, val in (("Origin", "0 0 0"), ("Spacing", "1 1 1")):
            want += ['    <DataStructure', f'       Name="{what}"', '       DataType="Double"', '       Dimensions="3"',
                     '       Format="XML">', f'       {val}', '    </DataStructure>']
        want += ['    </Geometry>']
        for nm in names:
            want += [f'      <Attribute Center="Node" Name="{nm}">', '        <DataStructure', '           DataType="Double"',
                     f'           Dimensions="{nz} {ny} {nx}"', '           Format="HDF">', f'           {base}.h5:/{nm}',
                     '        </DataStructure>', '      </Attribute>']
        want += ['   </Grid>']
    want += ['   </Grid>', ' </Domain>', '</Xdmf>']
    assert got == want
    # single-step flavour: its own file name, outputs iStep and iStep+1 (:154-158)
    ppk.write_xdmf_from_ini(ini, 5, single_step=True)
    one = open(tmp_path / "ot3d_0000005.xmf").read()
    assert one.count('GridType="Uniform"') == 2 and "ot3d_0000005.h5:/bz" in one and "ot3d_0000006.h5:/rho" in one

    U = np.zeros((8, nz + 6, ny + 6, nx + 6))
    if ppk.hdf5_available():
        ppk.save_data_from_ini(ini, U, 0)
        assert (tmp_path / "ot3d_0000000.h5").stat().st_size > 8 * nx * ny * nz * 8
    else:
        with pytest.raises(ppk.PpkError, match="10002"):  # PPK_ERR_UNSUPPORTED, after the .vti was written
            ppk.save_data_from_ini(ini, U, 0)
        assert (tmp_path / "ot3d_0000000.vti").exists()
        # a restart that cannot be honoured stops the program instead of starting something else
        import subprocess
        exe = os.path.join(ROOT, "ppkmhd_b200", "bin", "ppkMHD_b200")
        open(tmp_path / "r.ini", "w").write(ini.replace("[mesh]", "restart_enabled=true\nrestart_filename=ot3d_0000000.h5\n[mesh]"))
        r = subprocess.run([exe, "r.ini"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode != 0 and ("restart_enabled" in r.stderr or "no CUDA device" in r.stderr), r.stderr[-500:]


def test_solverbase_members_besides_the_time_loop(tmp_path):
    """SolverBase members a reference-side caller may use besides the time loop (src/shared/SolverBase.h:137, 166-198):
    save_data_debug writes outputPrefix_<debug_name>_%07d.vti next to save_data's file (IO_VTK.cpp:278-281), with the same
    payload; make_boundaries_serial / _mpi reach the solver's ghost fill; read_restart_file is the reference's empty TODO;
    load_data (restart) stops the program with the reason when it cannot be honoured. A C++ harness with a dummy solver
    (tests/host_harness/solverbase_members.cpp), linked against the in-tree library; no GPU involved."""
    import subprocess
    from oracle import oracle as O

    host = os.path.join(ROOT, "ppkmhd_b200", "host")
    libdir = os.path.join(ROOT, "ppkmhd_b200", "lib")
    exe = str(tmp_path / "harness")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-I", host, "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "host_harness", "solverbase_members.cpp"), "-L", libdir, "-lppkmhd_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    ini = O.make_ini("orszag_tang", (10, 6, 4)).replace("outputPrefix=run", f"outputDir={tmp_path}\noutputPrefix=dbg")
    open(tmp_path / "a.ini", "w").write(ini)
    r = subprocess.run([exe, "a.ini", "debug"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and "fills=2" in r.stdout, r.stdout + r.stderr
    plain, dbg = tmp_path / "dbg_0000003.vti", tmp_path / "dbg_afterBC_0000003.vti"
    assert plain.exists() and dbg.exists()
    assert plain.read_bytes() == dbg.read_bytes()
    import ppkmhd_b200 as ppk
    if not ppk.hdf5_available():
        open(tmp_path / "r.ini", "w").write(ini.replace("[mesh]", "restart_enabled=true\nrestart_filename=dbg_0000003.h5\n[mesh]"))
        r = subprocess.run([exe, "r.ini", "load"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode != 0 and "load_data:" in r.stderr and "returned" not in r.stdout, r.stdout + r.stderr

