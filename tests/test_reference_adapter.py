"""The reference-side binding of INTEGRATION.md section 2, compiled for real: oracle/_ref/ppkMHD_b200adapter is the
UNMODIFIED reference (its own main.cpp, SolverBase, HydroParams, ConfigMap, init functors, IO_VTK; Kokkos-OpenMP host
build) with one translation unit replaced (SolverFactory -> oracle/ref_build/adapter/SolverFactory_b200.cpp) so that
"MHD_Muscl_3D" creates oracle/ref_build/adapter/SolverMHDMusclB200.h, which drives libppkmhd_b200.so through the C ABI.

GPU test: that binary and the unmodified reference binary (oracle/_ref/ppkMHD) run the same .ini; every .vti they write
must be identical byte for byte, and so must the time-step log lines."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

ADAPTER = os.path.join(ROOT, "oracle", "_ref", "ppkMHD_b200adapter")
REFERENCE = os.path.join(ROOT, "oracle", "_ref", "ppkMHD")


def _run(exe, ini, tmp):
    open(os.path.join(tmp, "run.ini"), "w").write(ini)
    r = subprocess.run([exe, "run.ini"], cwd=tmp, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="4"))
    return r, sorted(f for f in os.listdir(tmp) if f.endswith(".vti"))


def test_adapter_binary_links_the_c_abi_library():
    """CPU: the adapter was built against the reference's headers and resolves libppkmhd_b200.so (no compute call)."""
    if not os.path.exists(ADAPTER):
        pytest.skip("oracle/_ref/ppkMHD_b200adapter not built (needs /root/reference: __graft_entry__.build())")
    out = subprocess.run(["ldd", ADAPTER], capture_output=True, text=True, check=True).stdout
    line = [ln for ln in out.splitlines() if "libppkmhd_b200.so" in ln]
    assert line and "not found" not in line[0], out
    syms = subprocess.run(["nm", "-D", "--undefined-only", ADAPTER], capture_output=True, text=True, check=True).stdout
    for name in ("ppk_mhd3d_create", "ppk_mhd3d_upload", "ppk_mhd3d_step", "ppk_mhd3d_download", "ppk_mhd3d_compute_dt",
                 "ppk_mhd3d_make_boundaries", "ppk_mhd3d_set_time", "ppk_mhd3d_get_time", "ppk_mhd3d_destroy"):
        assert name in syms, name


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ot_16x12x8", "blast_mixedbc_12x12x8", "fieldloop_24x12x12", "kh_sine_12x10x16", "wave_fast_16x8x8"])
def test_adapter_vti_equals_reference_vti(case):
    if not (os.path.exists(ADAPTER) and os.path.exists(REFERENCE)):
        pytest.skip("oracle/_ref binaries did not travel to this box")
    ini = str(np.load(f"{GOLDEN}/{case}.npz")["ini"])
    with tempfile.TemporaryDirectory() as ta, tempfile.TemporaryDirectory() as tr:
        ra, fa = _run(ADAPTER, ini, ta)
        assert ra.returncode == 0, ra.stdout[-2000:] + ra.stderr[-2000:]
        rr, fr = _run(REFERENCE, ini, tr)
        assert rr.returncode == 0, rr.stderr[-2000:]
        assert fa == fr and len(fa) >= 2, (fa, fr)
        for f in fa:
            assert open(os.path.join(ta, f), "rb").read() == open(os.path.join(tr, f), "rb").read(), f
        steps = lambda out: [ln for ln in out.splitlines() if ln.startswith("time step=") or ln.startswith("final time")]
        assert steps(ra.stdout) == steps(rr.stdout)
        assert "libppkmhd_b200" in ra.stdout
