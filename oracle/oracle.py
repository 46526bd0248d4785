"""ctypes front-end of the plain-C oracle (oracle/mhd3d_oracle.c)  --  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (ppkmhd_b200/) never does.

It also knows how to drive the UNMODIFIED reference binary (oracle/_ref/ppkMHD, built by
oracle/ref_build/Makefile from /root/reference) and to read the reference's binary .vti output
(src/utils/io/IO_VTK.cpp:211-408), which is the wire format parity is checked through.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ppkMHD")
VAR_NAMES = ["rho", "energy", "rho_vx", "rho_vy", "rho_vz", "bx", "by", "bz"]  # SolverBase.cpp:48-55
ID, IP, IU, IV, IW, IA, IB, IC = range(8)


def build(force: bool = False) -> str:
    # always ask make: the Makefile tracks mhd3d_oracle.c, mhd2d_oracle.c and mhd3d_oracle.h (a no-op when up to date)
    subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


class OrcParams(C.Structure):
    _fields_ = (
        [(n, C.c_int) for n in ("nx", "ny", "nz", "gw", "isize", "jsize", "ksize")]
        + [(n, C.c_double) for n in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "dx", "dy", "dz")]
        + [("bc", C.c_int * 6)]
        + [(n, C.c_double) for n in ("gamma0", "cfl", "slope_type", "smallr", "smallc", "smallp")]
        + [(n, C.c_int) for n in ("mx", "my", "mz", "px", "py", "pz")]
        + [("riemann", C.c_int)]
    )

    @property
    def shape(self):
        return (8, self.ksize, self.jsize, self.isize)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        pp = C.POINTER(OrcParams)
        _lib.orc_parse_float.restype = C.c_double
        _lib.orc_parse_float.argtypes = [C.c_char_p, C.c_double]
        _lib.orc_params_finalize.argtypes = [pp]
        _lib.orc_init_orszag_tang.argtypes = [pp, C.c_double, dp]
        _lib.orc_init_blast.argtypes = [pp] + [C.c_double] * 8 + [dp]
        _lib.orc_init_field_loop.argtypes = [pp] + [C.c_double] * 4 + [dp]
        _lib.orc_init_implode.argtypes = [pp, dp, dp, C.c_int, dp]
        _lib.orc_init_kelvin_helmholtz.argtypes = [pp] + [C.c_double] * 5 + [C.c_int, C.c_double, C.c_double, C.c_int, dp]
        _lib.orc_init_rotor.argtypes = [pp] + [C.c_double] * 5 + [dp]
        _lib.orc_init_wave.argtypes = [pp, C.c_double, C.c_int, dp]
        # 2-D path (mhd2d_oracle.c)
        _lib.orc2d_params_finalize.argtypes = [pp]
        _lib.orc2d_init_orszag_tang.argtypes = [pp, dp]
        _lib.orc2d_make_boundaries.argtypes = [pp, dp]
        _lib.orc2d_init_blast.argtypes = [pp] + [C.c_double] * 7 + [dp]
        _lib.orc2d_init_implode.argtypes = [pp, dp, dp, C.c_int, dp]
        _lib.orc2d_init_rotor.argtypes = [pp] + [C.c_double] * 5 + [dp]
        _lib.orc2d_init_field_loop.argtypes = [pp] + [C.c_double] * 4 + [dp]
        _lib.orc2d_init_kelvin_helmholtz.argtypes = [pp] + [C.c_double] * 5 + [C.c_int, C.c_double, C.c_double, C.c_int, dp]
        _lib.orc2d_step.restype = C.c_double
        _lib.orc2d_step.argtypes = [pp, dp, dp, dp, C.c_double, C.c_double]
        _lib.orc_make_boundary.argtypes = [pp, dp, C.c_int]
        _lib.orc_make_boundaries.argtypes = [pp, dp]
        _lib.orc_convert_to_primitives.argtypes = [pp, dp, dp]
        _lib.orc_compute_inv_dt.restype = C.c_double
        _lib.orc_compute_inv_dt.argtypes = [pp, dp]
        _lib.orc_compute_dt_local.restype = C.c_double
        _lib.orc_compute_dt_local.argtypes = [pp, dp]
        _lib.orc_scratch_create.restype = C.c_void_p
        _lib.orc_scratch_create.argtypes = [pp]
        _lib.orc_scratch_destroy.argtypes = [C.c_void_p]
        _lib.orc_godunov_v0.argtypes = [pp, dp, dp, dp, C.c_void_p, C.c_double]
        _lib.orc_step.restype = C.c_double
        _lib.orc_step.argtypes = [pp, dp, dp, dp, C.c_void_p, C.c_double, C.c_double]
        _lib.orc_scratch_array.restype = dp
        _lib.orc_scratch_array.argtypes = [C.c_void_p, C.c_char_p]
        _lib.orc_diagnostics.argtypes = [pp, dp, dp, dp]
    return _lib


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def parse_float(text: str, default: float = 0.0) -> float:
    """ConfigMap::getFloat semantics (float precision), src/utils/config/ConfigMap.cpp:37-46."""
    return lib().orc_parse_float(text.encode(), default)


# ---------------------------------------------------------------------------------------------
# minimal ini handling (independent of the product's C++ ConfigMap)
# ---------------------------------------------------------------------------------------------
def parse_ini(text: str) -> dict:
    """inih semantics used by the reference (src/utils/config/inih/ini.cpp): keys lower-cased
    'section.name' (INIReader.cpp:99-106), ';' after whitespace starts a comment, '#'/';' lines."""
    out, section = {}, ""
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line[0] in ";#":
            continue
        if line[0] == "[":
            section = line[1 : line.index("]")]
            continue
        if "=" not in line:
            continue
        name, value = line.split("=", 1)
        value = re.split(r"\s;", value, maxsplit=1)[0]
        out[(section + "." + name.strip()).lower()] = value.strip()
    return out


class Config:
    def __init__(self, text: str):
        self.text = text
        self.kv = parse_ini(text)

    def s(self, sec, name, default=""):
        return self.kv.get((sec + "." + name).lower(), default)

    def i(self, sec, name, default=0):
        v = self.s(sec, name, "")
        try:
            return int(v, 0)
        except ValueError:
            return default

    def b(self, sec, name, default=False):
        """ConfigMap::getBool (src/utils/config/ConfigMap.cpp:64-82)"""
        v = self.s(sec, name, "")
        if v in ("1", "yes", "true", "on"):
            return True
        if v in ("0", "no", "false", "off"):
            return False
        return default

    def f(self, sec, name, default=0.0):
        v = self.s(sec, name, "")
        return parse_float(v, default) if v else float(np.float32(default))


def params_from_config(cfg: Config, rank_pos=(0, 0, 0)) -> OrcParams:
    """HydroParams::setup for MHD_Muscl_3D (src/shared/HydroParams.cpp:28-217, 223-414)."""
    p = OrcParams()
    p.nx, p.ny, p.nz, p.gw = cfg.i("mesh", "nx", 1), cfg.i("mesh", "ny", 1), cfg.i("mesh", "nz", 1), 3
    p.xmin, p.ymin, p.zmin = (cfg.f("mesh", k, 0.0) for k in ("xmin", "ymin", "zmin"))
    p.xmax, p.ymax, p.zmax = (cfg.f("mesh", k, 1.0) for k in ("xmax", "ymax", "zmax"))
    p.mx, p.my, p.mz = cfg.i("mpi", "mx", 1), cfg.i("mpi", "my", 1), cfg.i("mpi", "mz", 1)
    p.px, p.py, p.pz = rank_pos
    names = ["xmin", "xmax", "ymin", "ymax", "zmin", "zmax"]
    m = [p.mx, p.mx, p.my, p.my, p.mz, p.mz]
    pos = [p.px, p.px, p.py, p.py, p.pz, p.pz]
    for f, nm in enumerate(names):
        bc = cfg.i("mesh", "boundary_type_" + nm, 1)
        outer = (pos[f] == 0) if f % 2 == 0 else (pos[f] == m[f] - 1)
        p.bc[f] = bc if outer else 4  # BC_COPY inside (HydroParams.cpp:300-351)
    p.gamma0 = cfg.f("hydro", "gamma0", 1.4)
    p.cfl = cfg.f("hydro", "cfl", 0.5)
    p.slope_type = cfg.f("hydro", "slope_type", 1.0)
    p.smallc = cfg.f("hydro", "smallc", 1e-10)
    p.smallr = cfg.f("hydro", "smallr", 1e-10)
    # HydroParams.cpp:175-198 (enum RiemannSolverType: approx 0, llf 1, hll 2, hllc 3, hlld 4)
    p.riemann = {"approx": 0, "llf": 1, "hll": 2, "hllc": 3, "hlld": 4}.get(cfg.s("hydro", "riemann", "approx"), 0)
    lib().orc_params_finalize(C.byref(p))
    return p


def init_problem(p: OrcParams, cfg: Config) -> np.ndarray:
    """SolverMHDMuscl::init dispatch (src/muscl/SolverMHDMuscl.h:653-713); unknown -> Orszag-Tang."""
    U = np.zeros(p.shape, dtype=np.float64)
    problem = cfg.s("hydro", "problem", "unknown")
    L = lib()
    if problem == "blast":
        # src/shared/problems/BlastParams.h:24-40 (defaults pass through a float argument)
        f32 = lambda x: float(np.float32(x))
        xmin, xmax = cfg.f("mesh", "xmin", 0.0), cfg.f("mesh", "xmax", 1.0)
        ymin, ymax = cfg.f("mesh", "ymin", 0.0), cfg.f("mesh", "ymax", 1.0)
        zmin, zmax = cfg.f("mesh", "zmin", 0.0), cfg.f("mesh", "zmax", 1.0)
        L.orc_init_blast(
            C.byref(p),
            cfg.f("blast", "radius", f32((xmin + xmax) / 2.0 / 10)),
            cfg.f("blast", "center_x", f32((xmin + xmax) / 2)),
            cfg.f("blast", "center_y", f32((ymin + ymax) / 2)),
            cfg.f("blast", "center_z", f32((zmin + zmax) / 2)),
            cfg.f("blast", "density_in", 1.0),
            cfg.f("blast", "density_out", 1.2),
            cfg.f("blast", "pressure_in", 10.0),
            cfg.f("blast", "pressure_out", 0.1),
            _dp(U),
        )
    elif problem in ("field_loop", "field loop"):
        L.orc_init_field_loop(
            C.byref(p),
            cfg.f("FieldLoop", "radius", 1.0),
            cfg.f("FieldLoop", "density_in", 1.0),
            cfg.f("FieldLoop", "amplitude", 1.0),
            cfg.f("FieldLoop", "vflow", 1.0),
            _dp(U),
        )
    elif problem == "implode":
        # src/shared/problems/ImplodeParams.h:36-58
        names = ("density", "pressure", "vx", "vy", "vz", "Bx", "By", "Bz")
        outer = np.array([cfg.f("implode", n + "_outer", d) for n, d in zip(names, (1.0, 1.0, 0, 0, 0, 0, 0, 0))])
        inner = np.array([cfg.f("implode", n + "_inner", d) for n, d in zip(names, (0.125, 0.14, 0, 0, 0, 0, 0, 0))])
        L.orc_init_implode(C.byref(p), _dp(outer), _dp(inner), cfg.i("implode", "shape_region", 0), _dp(U))
    elif problem == "kelvin_helmholtz":
        # src/shared/problems/KHParams.h:37-88
        if cfg.b("KH", "perturbation_rand", False):
            raise ValueError("perturbation_rand is not reproducible in the reference (per-thread Kokkos random pool)")
        rob, sine = cfg.b("KH", "perturbation_sine_robertson", True), cfg.b("KH", "perturbation_sine", False)
        if rob or sine:
            L.orc_init_kelvin_helmholtz(
                C.byref(p), cfg.f("KH", "d_in", 1.0), cfg.f("KH", "d_out", 1.0), cfg.f("KH", "pressure", 10.0),
                cfg.f("KH", "vflow_in", -0.5), cfg.f("KH", "vflow_out", 0.5), cfg.i("KH", "mode", 2),
                cfg.f("KH", "w0", 0.1), cfg.f("KH", "delta", 0.03), 1 if rob else 0, _dp(U))
    elif problem == "rotor":
        # src/shared/problems/RotorParams.h:17-24
        L.orc_init_rotor(C.byref(p), cfg.f("rotor", "r0", 0.1), cfg.f("rotor", "r1", 0.115), cfg.f("rotor", "u0", 2.0),
                         cfg.f("rotor", "p0", 1.0), cfg.f("rotor", "b0", 5.0 / np.sqrt(4 * np.pi)), _dp(U))
    elif problem == "wave":
        # src/shared/problems/WaveParams.h:47-49 (gamma0 and the mesh bounds are the ones already in p)
        if L.orc_init_wave(C.byref(p), cfg.f("wave", "amplitude", 1.0e-6), cfg.i("wave", "type", 0), _dp(U)) != 0:
            raise ValueError("wave type not implemented (the reference aborts)")
    else:
        L.orc_init_orszag_tang(C.byref(p), cfg.f("OrszagTang", "kt", 0.0), _dp(U))
    return U


class Oracle:
    """Single-rank driver mirroring main.cpp:133-166 + SolverBase::next_iteration for the v0 path."""

    def __init__(self, ini_text: str, rank_pos=(0, 0, 0)):
        self.cfg = Config(ini_text)
        self.p = params_from_config(self.cfg, rank_pos)
        self.t_end = self.cfg.f("run", "tEnd", 0.0)
        self.nstepmax = self.cfg.i("run", "nstepmax", 1000)
        self.t = self.cfg.f("run", "tCurrent", 0.0)
        self.iteration = 0
        self.dt = self.t_end
        self.U = init_problem(self.p, self.cfg)
        self.U2 = np.zeros_like(self.U)
        self.Q = np.zeros_like(self.U)
        self.scratch = lib().orc_scratch_create(C.byref(self.p))
        # constructor sequence, src/muscl/SolverMHDMuscl.h:390-402
        if self.p.mx * self.p.my * self.p.mz == 1:
            lib().orc_make_boundaries(C.byref(self.p), _dp(self.U))
            self.U2[...] = self.U

    def __del__(self):
        try:
            lib().orc_scratch_destroy(self.scratch)
        except Exception:
            pass

    @property
    def current(self) -> np.ndarray:
        return self.U if self.iteration % 2 == 0 else self.U2

    def finished(self) -> bool:  # SolverBase.cpp:196-201
        return self.t >= (self.t_end - 1e-14) or self.iteration >= self.nstepmax

    def step(self) -> float:
        a, b = (self.U, self.U2) if self.iteration % 2 == 0 else (self.U2, self.U)
        self.dt = lib().orc_step(C.byref(self.p), _dp(a), _dp(b), _dp(self.Q), self.scratch, self.t, self.t_end)
        self.iteration += 1
        self.t += self.dt
        return self.dt

    def run(self, nsteps=None):
        n = 0
        while not self.finished() and (nsteps is None or n < nsteps):
            self.step()
            n += 1
        return self

    def interior(self, U=None) -> np.ndarray:
        U = self.current if U is None else U
        g = self.p.gw
        return U[:, g:-g, g:-g, g:-g]

    def scratch_array(self, name: str, ncomp: int) -> np.ndarray:
        ptr = lib().orc_scratch_array(self.scratch, name.encode())
        n = self.p.isize * self.p.jsize * self.p.ksize
        return np.ctypeslib.as_array(ptr, shape=(ncomp * n,)).reshape(ncomp, self.p.ksize, self.p.jsize, self.p.isize)

    def diagnostics(self, U=None):
        U = np.ascontiguousarray(self.current if U is None else U)
        V = U.copy()
        lib().orc_make_boundaries(C.byref(self.p), _dp(V))
        sums = np.zeros(8)
        m = C.c_double(0)
        lib().orc_diagnostics(C.byref(self.p), _dp(V), _dp(sums), C.byref(m))
        return sums, m.value


def init_problem_2d(p: OrcParams, cfg: Config, U: np.ndarray) -> None:
    """SolverMHDMuscl<2>::init dispatch (src/muscl/SolverMHDMuscl.h:653-713) with the 2-D functors of MHDInitFunctors2D.h;
    the reference's 2-D wave functor is empty (ValueError); an unknown name falls back to Orszag-Tang like the reference."""
    L = lib()
    problem = cfg.s("hydro", "problem", "unknown")
    f32 = lambda x: float(np.float32(x))
    if problem == "blast":
        xmin, xmax = cfg.f("mesh", "xmin", 0.0), cfg.f("mesh", "xmax", 1.0)
        ymin, ymax = cfg.f("mesh", "ymin", 0.0), cfg.f("mesh", "ymax", 1.0)
        L.orc2d_init_blast(C.byref(p), cfg.f("blast", "radius", f32((xmin + xmax) / 2.0 / 10)),
                           cfg.f("blast", "center_x", f32((xmin + xmax) / 2)), cfg.f("blast", "center_y", f32((ymin + ymax) / 2)),
                           cfg.f("blast", "density_in", 1.0), cfg.f("blast", "density_out", 1.2),
                           cfg.f("blast", "pressure_in", 10.0), cfg.f("blast", "pressure_out", 0.1), _dp(U))
    elif problem == "rotor":
        L.orc2d_init_rotor(C.byref(p), cfg.f("rotor", "r0", 0.1), cfg.f("rotor", "r1", 0.115), cfg.f("rotor", "u0", 2.0),
                           cfg.f("rotor", "p0", 1.0), cfg.f("rotor", "b0", 5.0 / np.sqrt(4 * np.pi)), _dp(U))
    elif problem in ("field_loop", "field loop"):
        L.orc2d_init_field_loop(C.byref(p), cfg.f("FieldLoop", "radius", 1.0), cfg.f("FieldLoop", "density_in", 1.0),
                                cfg.f("FieldLoop", "amplitude", 1.0), cfg.f("FieldLoop", "vflow", 1.0), _dp(U))
    elif problem == "kelvin_helmholtz":
        if cfg.b("KH", "perturbation_rand", False):
            raise ValueError("perturbation_rand is not reproducible in the reference (per-thread Kokkos random pool)")
        rob, sine = cfg.b("KH", "perturbation_sine_robertson", True), cfg.b("KH", "perturbation_sine", False)
        if rob or sine:
            L.orc2d_init_kelvin_helmholtz(
                C.byref(p), cfg.f("KH", "d_in", 1.0), cfg.f("KH", "d_out", 1.0), cfg.f("KH", "pressure", 10.0),
                cfg.f("KH", "vflow_in", -0.5), cfg.f("KH", "vflow_out", 0.5), cfg.i("KH", "mode", 2),
                cfg.f("KH", "w0", 0.1), cfg.f("KH", "delta", 0.03), 1 if rob else 0, _dp(U))
    elif problem == "implode":
        names = ("density", "pressure", "vx", "vy", "vz", "Bx", "By", "Bz")
        outer = np.array([cfg.f("implode", n + "_outer", d) for n, d in zip(names, (1.0, 1.0, 0, 0, 0, 0, 0, 0))])
        inner = np.array([cfg.f("implode", n + "_inner", d) for n, d in zip(names, (0.125, 0.14, 0, 0, 0, 0, 0, 0))])
        L.orc2d_init_implode(C.byref(p), _dp(outer), _dp(inner), cfg.i("implode", "shape_region", 0), _dp(U))
    elif problem == "wave":
        raise ValueError("InitWaveFunctor2D_MHD is an empty functor in the reference (MHDInitFunctors2D.h:950-980): no 2-D wave")
    else:
        L.orc2d_init_orszag_tang(C.byref(p), _dp(U))


class Oracle2D:
    """The 2-D path (MHD_Muscl_2D, implementationVersion 0): oracle only, no CUDA counterpart yet.
    Arrays are (8, jsize, isize)."""

    def __init__(self, ini_text: str):
        cfg = self.cfg = Config(ini_text)
        p = self.p = OrcParams()
        p.nx, p.ny, p.nz, p.gw = cfg.i("mesh", "nx", 1), cfg.i("mesh", "ny", 1), 1, 3
        p.xmin, p.ymin, p.zmin = (cfg.f("mesh", k, 0.0) for k in ("xmin", "ymin", "zmin"))
        p.xmax, p.ymax, p.zmax = (cfg.f("mesh", k, 1.0) for k in ("xmax", "ymax", "zmax"))
        p.mx = p.my = p.mz = 1
        for f, nm in enumerate(("xmin", "xmax", "ymin", "ymax")):
            p.bc[f] = cfg.i("mesh", "boundary_type_" + nm, 1)
        p.gamma0, p.cfl = cfg.f("hydro", "gamma0", 1.4), cfg.f("hydro", "cfl", 0.5)
        p.slope_type = cfg.f("hydro", "slope_type", 1.0)
        p.smallc, p.smallr = cfg.f("hydro", "smallc", 1e-10), cfg.f("hydro", "smallr", 1e-10)
        p.riemann = {"approx": 0, "llf": 1, "hll": 2, "hllc": 3, "hlld": 4}.get(cfg.s("hydro", "riemann", "approx"), 0)
        lib().orc2d_params_finalize(C.byref(p))
        self.t_end = cfg.f("run", "tEnd", 0.0)
        self.nstepmax = cfg.i("run", "nstepmax", 1000)
        self.t, self.iteration, self.dt = 0.0, 0, self.t_end
        self.U = np.zeros((8, p.jsize, p.isize))
        init_problem_2d(p, cfg, self.U)
        lib().orc2d_make_boundaries(C.byref(p), _dp(self.U))  # constructor sequence, SolverMHDMuscl.h:390-402
        self.U2 = self.U.copy()
        self.Q = np.zeros_like(self.U)

    @property
    def current(self):
        return self.U if self.iteration % 2 == 0 else self.U2

    def finished(self):
        return self.t >= (self.t_end - 1e-14) or self.iteration >= self.nstepmax

    def step(self):
        a, b = (self.U, self.U2) if self.iteration % 2 == 0 else (self.U2, self.U)
        self.dt = lib().orc2d_step(C.byref(self.p), _dp(a), _dp(b), _dp(self.Q), self.t, self.t_end)
        self.iteration += 1
        self.t += self.dt
        return self.dt

    def run(self, nsteps=None):
        n = 0
        while not self.finished() and (nsteps is None or n < nsteps):
            self.step()
            n += 1
        return self

    def interior(self):
        g = self.p.gw
        return self.current[:, g:-g, g:-g]


# ---------------------------------------------------------------------------------------------
# reference binary + VTI
# ---------------------------------------------------------------------------------------------
def read_vti(path: str) -> np.ndarray:
    """Binary appended-raw .vti written by save_VTK_3D (IO_VTK.cpp:359-399): returns (8,nz,ny,nx)."""
    with open(path, "rb") as f:
        blob = f.read()
    head_end = blob.index(b'<AppendedData encoding="raw">')
    header = blob[:head_end].decode()
    ext = re.search(r'WholeExtent="0 (\d+) 0 (\d+) 0 (\d+)"', header)
    nx, ny, nz = (int(ext.group(i)) for i in (1, 2, 3))
    nz = max(nz, 1)  # 2-D files carry "0 nx 0 ny 0 0"
    names = re.findall(r'Name="([^"]+)"', header)
    pos = blob.index(b"_", head_end) + 1
    out = {}
    for nm in names:
        nbytes = int(np.frombuffer(blob, dtype="<u8", count=1, offset=pos)[0])
        pos += 8
        out[nm] = np.frombuffer(blob, dtype="<f8", count=nbytes // 8, offset=pos).reshape(nz, ny, nx)
        pos += nbytes
    return np.stack([out[n] for n in VAR_NAMES])


def have_reference() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_reference(ini_text: str, threads: int | None = None, workdir: str | None = None, keep=False):
    """Run the unmodified reference on `ini_text`; returns (stdout, [vti arrays in output order])."""
    if not have_reference():
        raise RuntimeError("oracle/_ref/ppkMHD is absent: run `make -C oracle ref` where /root/reference exists")
    tmp = workdir or tempfile.mkdtemp(prefix="ppkref_")
    with open(os.path.join(tmp, "run.ini"), "w") as f:
        f.write(ini_text)
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads or os.cpu_count() or 1)
    env["OMP_PROC_BIND"] = "spread"
    env["OMP_PLACES"] = "threads"
    res = subprocess.run([REF_BIN, "run.ini"], cwd=tmp, env=env, capture_output=True, text=True, check=True)
    files = sorted(f for f in os.listdir(tmp) if f.endswith(".vti"))
    states = [read_vti(os.path.join(tmp, f)) for f in files]
    if not keep and workdir is None:
        for f in os.listdir(tmp):
            os.remove(os.path.join(tmp, f))
        os.rmdir(tmp)
    return res.stdout, states


def make_ini(problem="orszag_tang", n=(32, 32, 32), nstepmax=5, tend=1.0, noutput=1, bounds=None, bc=3,
             cfl=0.8, extra="", mz=1, prefix="run", nlog=10, riemann="hlld", mx=1, my=1) -> str:
    """The ini family of SURVEY 8(d): gamma0=1.666 cfl=0.8 slope_type=2 hlld smallr=smallc=1e-8, v0."""
    b = bounds or (0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    bcs = bc if isinstance(bc, (list, tuple)) else [bc] * 6
    names = ["xmin", "xmax", "ymin", "ymax", "zmin", "zmax"]
    bc_txt = "\n".join(f"boundary_type_{nm}={v}" for nm, v in zip(names, bcs))
    return f"""[run]
solver_name=MHD_Muscl_3D
tEnd={tend}
nStepmax={nstepmax}
nOutput={noutput}
nlog={nlog}
[mesh]
nx={n[0]}
ny={n[1]}
nz={n[2]}
xmin={b[0]}
xmax={b[1]}
ymin={b[2]}
ymax={b[3]}
zmin={b[4]}
zmax={b[5]}
{bc_txt}
[hydro]
gamma0=1.666
cfl={cfl}
niter_riemann=10
iorder=2
slope_type=2
problem={problem}
riemann={riemann}
smallr=1e-8
smallc=1e-8
[mpi]
mx={mx}
my={my}
mz={mz}
[output]
outputPrefix={prefix}
outputVtkAscii=false
[other]
implementationVersion=0
{extra}
"""
