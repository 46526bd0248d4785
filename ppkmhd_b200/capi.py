"""ctypes binding of include/ppkmhd_b200.h (+ the host-layer entry points of
include/ppkmhd_b200_host.h).  The reference-side stub a maintainer would write is shown in
INTEGRATION.md; this is the same thing for Python callers (tests, bench.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    # PPKMHD_B200_LIB: an alternative build of the same library (kernel-tuning experiments)
    return os.environ.get("PPKMHD_B200_LIB") or os.path.join(HERE, "lib", "libppkmhd_b200.so")


def build_library(verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (ppkmhd_b200/Makefile); works without a GPU."""
    subprocess.check_call(["make", "-C", HERE, "-j4"], stdout=None if verbose else subprocess.DEVNULL)
    return lib_path()


class PpkError(RuntimeError):
    pass


class Params(C.Structure):
    """struct ppk_mhd3d_params"""

    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("ghost_width", C.c_int),
        ("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
        ("zmin", C.c_double), ("zmax", C.c_double),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("boundary_type", C.c_int * 6),
        ("gamma0", C.c_double), ("cfl", C.c_double), ("slope_type", C.c_double),
        ("smallr", C.c_double), ("smallc", C.c_double), ("smallp", C.c_double),
        ("riemann_solver", C.c_int), ("implementation_version", C.c_int),
        ("mx", C.c_int), ("my", C.c_int), ("mz", C.c_int),
        ("rank_x", C.c_int), ("rank_y", C.c_int), ("rank_z", C.c_int),
        ("device", C.c_int), ("exact_arithmetic", C.c_int),
    ]

    @property
    def shape(self):
        return (8, self.nz + 6, self.ny + 6, self.nx + 6)


class HaloMsg(C.Structure):
    """struct ppk_halo_msg"""

    _fields_ = [("peer", C.c_int), ("is_send", C.c_int), ("var", C.c_int), ("offset", C.c_longlong), ("count", C.c_longlong)]


class FaceMsg(C.Structure):
    """struct ppk_face_msg"""

    _fields_ = [("peer", C.c_int), ("is_send", C.c_int), ("hi_face", C.c_int), ("first_layer", C.c_int), ("count", C.c_longlong)]


_lib = None

# every symbol include/*.h declares (tests/test_host_layer.py checks the library exports them all)
EXPORTS = [
    "ppk_mhd3d_create", "ppk_mhd3d_destroy", "ppk_mhd3d_upload", "ppk_mhd3d_download", "ppk_mhd3d_download_async", "ppk_mhd3d_set_time",
    "ppk_mhd3d_get_time", "ppk_mhd3d_make_boundaries", "ppk_mhd3d_compute_dt", "ppk_mhd3d_step", "ppk_mhd3d_run",
    "ppk_mhd3d_synchronize", "ppk_mhd3d_diagnostics", "ppk_nccl_get_unique_id", "ppk_mhd3d_comm_init",
    "ppk_mhd3d_set_stream", "ppk_mhd3d_profile", "ppk_mhd3d_kernel_times", "ppk_mhd3d_kernel_timeline", "ppk_mhd3d_launch_count",
    "ppk_mhd3d_debug_array", "ppk_mhd3d_device_bytes", "ppk_last_error_string", "ppk_version_string",
    "ppk_mhd3d_halo_plan", "ppk_mhd3d_face_plan", "ppk_selftest_fastmath", "ppk_mhd3d_set_pipeline",
    "ppk_mhd3d_get_pipeline", "ppk_mhd3d_stage_upload", "ppk_mhd3d_stage_swap", "ppk_mhd3d_stage_download",
    "ppk_params_from_ini", "ppk_init_condition_from_ini", "ppk_run_ini", "ppk_save_data_from_ini", "ppk_load_data_from_ini", "ppk_hdf5_available", "ppk_write_xdmf_from_ini", "ppk_init_condition_2d_from_ini",
    "ppk_mhd2d_create", "ppk_mhd2d_destroy", "ppk_mhd2d_upload", "ppk_mhd2d_download", "ppk_mhd2d_set_time", "ppk_mhd2d_get_time",
    "ppk_mhd2d_make_boundaries", "ppk_mhd2d_compute_dt", "ppk_mhd2d_step", "ppk_mhd2d_run", "ppk_mhd2d_synchronize", "ppk_mhd2d_launch_count",
]


def load_library():
    """Load the CUDA library. Fails loudly if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise PpkError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.ppk_mhd3d_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.ppk_mhd3d_destroy.argtypes = [vp]
    L.ppk_mhd3d_upload.argtypes = [vp, vp]
    L.ppk_mhd3d_download.argtypes = [vp, vp]
    L.ppk_mhd3d_download_async.argtypes = [vp, vp]
    L.ppk_mhd3d_stage_upload.argtypes = [vp, vp]
    L.ppk_mhd3d_stage_swap.argtypes = [vp]
    L.ppk_mhd3d_stage_download.argtypes = [vp, vp]
    L.ppk_mhd3d_set_time.argtypes = [vp, C.c_double, C.c_double, C.c_long]
    L.ppk_mhd3d_get_time.argtypes = [vp, dp, dp, C.POINTER(C.c_long)]
    L.ppk_mhd3d_make_boundaries.argtypes = [vp]
    L.ppk_mhd3d_compute_dt.argtypes = [vp, dp]
    L.ppk_mhd3d_step.argtypes = [vp]
    L.ppk_mhd3d_run.argtypes = [vp, C.c_int]
    L.ppk_mhd3d_synchronize.argtypes = [vp]
    L.ppk_mhd3d_diagnostics.argtypes = [vp, dp, dp]
    L.ppk_nccl_get_unique_id.argtypes = [vp]
    L.ppk_mhd3d_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.ppk_mhd3d_set_stream.argtypes = [vp, vp]
    L.ppk_mhd3d_profile.argtypes = [vp, C.c_int]
    L.ppk_mhd3d_set_pipeline.argtypes = [vp, C.c_int]
    L.ppk_mhd3d_get_pipeline.argtypes = [vp]
    L.ppk_mhd3d_kernel_times.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), dp, C.POINTER(C.c_long), C.c_int]
    L.ppk_mhd3d_launch_count.argtypes = [vp]
    L.ppk_mhd3d_launch_count.restype = C.c_long
    L.ppk_mhd3d_debug_array.argtypes = [vp, C.c_char_p, vp, ip]
    L.ppk_mhd3d_device_bytes.argtypes = [vp]
    L.ppk_mhd3d_device_bytes.restype = C.c_longlong
    L.ppk_mhd3d_halo_plan.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(HaloMsg)]
    L.ppk_mhd3d_face_plan.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(FaceMsg)]
    L.ppk_selftest_fastmath.argtypes = [C.c_int, vp, vp, vp, vp]
    L.ppk_last_error_string.restype = C.c_char_p
    L.ppk_version_string.restype = C.c_char_p
    L.ppk_params_from_ini.argtypes = [C.c_char_p, C.c_int, C.POINTER(Params), dp, ip]
    L.ppk_init_condition_from_ini.argtypes = [C.c_char_p, C.c_int, vp]
    L.ppk_save_data_from_ini.argtypes = [C.c_char_p, C.c_int, vp, C.c_int]
    L.ppk_write_xdmf_from_ini.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.ppk_load_data_from_ini.argtypes = [C.c_char_p, C.c_int, vp, C.POINTER(C.c_int), dp]
    L.ppk_run_ini.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.ppk_init_condition_2d_from_ini.argtypes = [C.c_char_p, vp]
    L.ppk_mhd2d_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    for name in ("destroy", "make_boundaries", "step", "synchronize"):
        getattr(L, "ppk_mhd2d_" + name).argtypes = [vp]
    L.ppk_mhd2d_upload.argtypes = [vp, vp]
    L.ppk_mhd2d_download.argtypes = [vp, vp]
    L.ppk_mhd2d_set_time.argtypes = [vp, C.c_double, C.c_double, C.c_long]
    L.ppk_mhd2d_get_time.argtypes = [vp, dp, dp, C.POINTER(C.c_long)]
    L.ppk_mhd2d_run.argtypes = [vp, C.c_int]
    L.ppk_mhd2d_compute_dt.argtypes = [vp, dp]
    L.ppk_mhd2d_launch_count.argtypes = [vp]
    L.ppk_mhd2d_launch_count.restype = C.c_long
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise PpkError(f"ppkmhd_b200 error {rc}: {load_library().ppk_last_error_string().decode()}")


def params_from_ini(ini_text: str, rank_z: int = 0, device: int = 0, exact: bool = True):
    """ConfigMap + HydroParams::setup of the C++ host layer. Returns (Params, t_end, nstepmax)."""
    L = load_library()
    p = Params()
    t_end = C.c_double(0)
    nstep = C.c_int(0)
    _check(L.ppk_params_from_ini(ini_text.encode(), rank_z, C.byref(p), C.byref(t_end), C.byref(nstep)))
    p.device = device
    p.exact_arithmetic = 1 if exact else 0
    return p, t_end.value, nstep.value


def init_condition_from_ini(ini_text: str, rank_z: int = 0) -> np.ndarray:
    """Host-side initial condition of the C++ host layer (SolverMHDMusclCuda3D::init)."""
    L = load_library()
    p, _, _ = params_from_ini(ini_text, rank_z)
    U = np.zeros(p.shape, dtype=np.float64)
    _check(L.ppk_init_condition_from_ini(ini_text.encode(), rank_z, U.ctypes.data))
    return U


def save_data_from_ini(ini_text: str, U: np.ndarray, i_step: int, rank_z: int = 0) -> None:
    """ppk_save_data_from_ini: the host layer's SolverBase::save_data for slab `rank_z` (.vti, or piece + .pvti); no GPU."""
    L = load_library()
    U = np.ascontiguousarray(U, dtype=np.float64)
    _check(L.ppk_save_data_from_ini(ini_text.encode(), rank_z, U.ctypes.data, i_step))


def load_data_from_ini(ini_text: str, U: np.ndarray, rank_z: int = 0):
    """ppk_load_data_from_ini: restart file ([run] restart_filename) -> U (in place), returns (output number, time)."""
    L = load_library()
    assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"]
    step, t = C.c_int(0), C.c_double(0.0)
    _check(L.ppk_load_data_from_ini(ini_text.encode(), rank_z, U.ctypes.data, C.byref(step), C.byref(t)))
    return step.value, t.value


def hdf5_available() -> bool:
    return bool(load_library().ppk_hdf5_available())


def write_xdmf_from_ini(ini_text: str, total_steps: int, single_step: bool = False) -> None:
    """ppk_write_xdmf_from_ini: the reference's Xdmf wrapper text (written into the current directory)."""
    _check(load_library().ppk_write_xdmf_from_ini(ini_text.encode(), total_steps, 1 if single_step else 0))


def init_condition_2d_from_ini(ini_text: str) -> np.ndarray:
    """Host-side Orszag-Tang initial condition of the 2-D path: (8, jsize, isize)."""
    L = load_library()
    p, _, _ = params_from_ini(ini_text)
    U = np.zeros((8, p.ny + 6, p.nx + 6), dtype=np.float64)
    _check(L.ppk_init_condition_2d_from_ini(ini_text.encode(), U.ctypes.data))
    return U


class Mhd2d:
    """The 2-D MHD solver (a ppk_mhd2d handle): MHD_Muscl_2D, implementationVersion 0, single GPU."""

    def __init__(self, params: Params):
        self.L = load_library()
        self.params = params
        self.shape = (8, params.ny + 6, params.nx + 6)
        self.h = C.c_void_p()
        _check(self.L.ppk_mhd2d_create(C.byref(params), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.L.ppk_mhd2d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, U: np.ndarray):
        assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"] and U.shape == self.shape
        _check(self.L.ppk_mhd2d_upload(self.h, U.ctypes.data))

    def download(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float64)
        _check(self.L.ppk_mhd2d_download(self.h, out.ctypes.data))
        return out

    def interior(self) -> np.ndarray:
        return self.download()[:, 3:-3, 3:-3]

    def set_time(self, t, t_end, iteration=0):
        _check(self.L.ppk_mhd2d_set_time(self.h, t, t_end, iteration))

    def get_time(self):
        t, dt, it = C.c_double(0), C.c_double(0), C.c_long(0)
        _check(self.L.ppk_mhd2d_get_time(self.h, C.byref(t), C.byref(dt), C.byref(it)))
        return t.value, dt.value, it.value

    def step(self):
        _check(self.L.ppk_mhd2d_step(self.h))

    def run(self, nsteps):
        _check(self.L.ppk_mhd2d_run(self.h, nsteps))

    def launch_count(self):
        return self.L.ppk_mhd2d_launch_count(self.h)


class Mhd3d:
    """One GPU slab of the 3-D MHD solver (a ppk_mhd3d handle)."""

    def __init__(self, params: Params):
        self.L = load_library()
        self.params = params
        self.h = C.c_void_p()
        _check(self.L.ppk_mhd3d_create(C.byref(params), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.L.ppk_mhd3d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def shape(self):
        return self.params.shape

    def upload(self, U):
        """U: numpy float64 (8,ksize,jsize,isize) C-contiguous, or the address of such a host buffer."""
        if isinstance(U, np.ndarray):
            assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"] and U.shape == self.shape
            _check(self.L.ppk_mhd3d_upload(self.h, U.ctypes.data))
        else:
            _check(self.L.ppk_mhd3d_upload(self.h, int(U)))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        if isinstance(out, np.ndarray):
            _check(self.L.ppk_mhd3d_download(self.h, out.ctypes.data))
        else:
            _check(self.L.ppk_mhd3d_download(self.h, int(out)))
        return out

    def download_async(self, out_ptr):
        """Enqueue the device-to-host copy into the PINNED buffer at address `out_ptr`; valid after synchronize()."""
        _check(self.L.ppk_mhd3d_download_async(self.h, int(out_ptr)))

    def stage_upload(self, in_ptr):
        """ppk_mhd3d_stage_upload: asynchronous H2D of a full state from the PINNED buffer at `in_ptr` into the staging array."""
        _check(self.L.ppk_mhd3d_stage_upload(self.h, int(in_ptr)))

    def stage_swap(self):
        _check(self.L.ppk_mhd3d_stage_swap(self.h))

    def stage_download(self, out_ptr):
        """ppk_mhd3d_stage_download: asynchronous D2H of the current array on the second copy stream; valid after synchronize()."""
        _check(self.L.ppk_mhd3d_stage_download(self.h, int(out_ptr)))

    def interior(self):
        return self.download()[:, 3:-3, 3:-3, 3:-3]

    def set_time(self, t, t_end, iteration=0):
        _check(self.L.ppk_mhd3d_set_time(self.h, t, t_end, iteration))

    def get_time(self):
        t, dt, it = C.c_double(), C.c_double(), C.c_long()
        _check(self.L.ppk_mhd3d_get_time(self.h, C.byref(t), C.byref(dt), C.byref(it)))
        return t.value, dt.value, it.value

    def make_boundaries(self):
        _check(self.L.ppk_mhd3d_make_boundaries(self.h))

    def compute_dt(self):
        dt = C.c_double()
        _check(self.L.ppk_mhd3d_compute_dt(self.h, C.byref(dt)))
        return dt.value

    def step(self):
        _check(self.L.ppk_mhd3d_step(self.h))

    def run(self, nsteps):
        _check(self.L.ppk_mhd3d_run(self.h, nsteps))

    def synchronize(self):
        _check(self.L.ppk_mhd3d_synchronize(self.h))

    def diagnostics(self):
        sums = np.zeros(8)
        m = C.c_double()
        _check(self.L.ppk_mhd3d_diagnostics(self.h, sums.ctypes.data_as(C.POINTER(C.c_double)), C.byref(m)))
        return sums, m.value

    def comm_init(self, unique_id: bytes, nranks: int, rank: int):
        buf = C.create_string_buffer(unique_id, 128)
        _check(self.L.ppk_mhd3d_comm_init(self.h, buf, nranks, rank))

    def set_stream(self, stream_ptr):
        _check(self.L.ppk_mhd3d_set_stream(self.h, stream_ptr))

    PIPELINES = {"unfused": 0, "fused": 1, "fused_split": 2, "streamed": 3, "tiled": 4, "ordered": 5}

    def set_pipeline(self, name: str):
        _check(self.L.ppk_mhd3d_set_pipeline(self.h, self.PIPELINES[name]))

    def pipeline(self) -> str:
        code = self.L.ppk_mhd3d_get_pipeline(self.h)
        return {v: k for k, v in self.PIPELINES.items()}.get(code, str(code))

    def profile(self, enable=True):
        _check(self.L.ppk_mhd3d_profile(self.h, 1 if enable else 0))

    def kernel_times(self, reset=False):
        n = 32
        names = (C.c_char_p * n)()
        ms = (C.c_double * n)()
        cnt = (C.c_long * n)()
        k = self.L.ppk_mhd3d_kernel_times(self.h, n, names, ms, cnt, 1 if reset else 0)
        if k < 0:
            raise PpkError(self.L.ppk_last_error_string().decode())
        return {names[i].decode(): (ms[i], cnt[i]) for i in range(k)}

    def kernel_timeline(self, capacity=4096):
        """ppk_mhd3d_kernel_timeline: [(name, on_comm_stream, start_ms, end_ms)] of the launches since profile(True)."""
        names = (C.c_char_p * capacity)()
        comm = (C.c_int * capacity)()
        t0 = (C.c_double * capacity)()
        t1 = (C.c_double * capacity)()
        self.L.ppk_mhd3d_kernel_timeline.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        k = self.L.ppk_mhd3d_kernel_timeline(self.h, capacity, names, comm, t0, t1)
        if k < 0:
            raise PpkError(self.L.ppk_last_error_string().decode())
        return [(names[i].decode(), bool(comm[i]), t0[i], t1[i]) for i in range(k)]

    def launch_count(self):
        return self.L.ppk_mhd3d_launch_count(self.h)

    def device_bytes(self):
        return self.L.ppk_mhd3d_device_bytes(self.h)

    def debug_array(self, name: str):
        nc = C.c_int()
        _check(self.L.ppk_mhd3d_debug_array(self.h, name.encode(), None, C.byref(nc)))
        out = np.empty((nc.value,) + self.shape[1:], dtype=np.float64)
        _check(self.L.ppk_mhd3d_debug_array(self.h, name.encode(), out.ctypes.data, C.byref(nc)))
        return out


def halo_plan(params: Params):
    """ppk_mhd3d_halo_plan: list of (peer, is_send, var, offset, count) of one z-halo exchange."""
    L = load_library()
    msgs = (HaloMsg * 32)()
    n = L.ppk_mhd3d_halo_plan(C.byref(params), 32, msgs)
    if n < 0:
        raise PpkError("ppk_mhd3d_halo_plan: bad arguments")
    return [(m.peer, bool(m.is_send), m.var, m.offset, m.count) for m in msgs[:n]]


def face_plan(params: Params, direction: int):
    """ppk_mhd3d_face_plan: list of (peer, is_send, hi_face, first_layer, count) of the packed x (0) or y (1) exchange."""
    L = load_library()
    msgs = (FaceMsg * 4)()
    n = L.ppk_mhd3d_face_plan(C.byref(params), direction, msgs)
    if n < 0:
        raise PpkError("ppk_mhd3d_face_plan: bad arguments")
    return [(m.peer, bool(m.is_send), bool(m.hi_face), m.first_layer, m.count) for m in msgs[:n]]


def selftest_fastmath(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = [np.empty_like(x) for _ in range(3)]
    _check(load_library().ppk_selftest_fastmath(x.size, x.ctypes.data, *(o.ctypes.data for o in out)))
    return out


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load_library().ppk_nccl_get_unique_id(buf))
    return buf.raw
