"""ppkmhd_b200 -- B200-native (sm_100a) MUSCL-Hancock + constrained-transport MHD step behind ppkMHD's
solver interface.  The product is the C-ABI shared library (include/ppkmhd_b200.h) and the C++ host
layer in ppkmhd_b200/host/; this Python package is only a thin ctypes binding used by the tests and
bench.py.  It never imports oracle/ and has no CPU fallback."""
from .capi import (  # noqa: F401
    Mhd2d,
    Mhd3d,
    Params,
    PpkError,
    build_library,
    halo_plan,
    face_plan,
    selftest_fastmath,
    init_condition_from_ini,
    save_data_from_ini,
    load_data_from_ini,
    hdf5_available,
    write_xdmf_from_ini,
    init_condition_2d_from_ini,
    lib_path,
    load_library,
    nccl_unique_id,
    params_from_ini,
)

__all__ = [
    "Mhd3d", "Mhd2d", "Params", "PpkError", "lib_path", "load_library", "build_library",
    "params_from_ini", "init_condition_from_ini", "save_data_from_ini", "load_data_from_ini", "hdf5_available", "write_xdmf_from_ini", "init_condition_2d_from_ini", "nccl_unique_id", "halo_plan", "face_plan", "selftest_fastmath",
]
