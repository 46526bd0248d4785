#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (k_producer, k_riemann_all, k_riemann_pers, k_trace_tma, k_face_copy is 4-GPU only)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
import ppkmhd_b200 as ppk
from oracle import oracle as O
ini = O.make_ini("orszag_tang", (64, 36, 10), nstepmax=2, extra="[OrszagTang]\nkt=1\n", tend=10.0)
ref = None
for pipeline in ("unfused", "ordered", "tiled"):
    for exact in (True, False):
        p, t_end, n = ppk.params_from_ini(ini, exact=exact)
        s = ppk.Mhd3d(p); s.set_pipeline(pipeline)
        s.upload(ppk.init_condition_from_ini(ini)); s.set_time(0.0, t_end, 0)
        s.run(2); u = s.interior().copy(); s.close()
        if exact:
            if ref is None: ref = u
            assert np.array_equal(u, ref), pipeline
print("sanitizer workload ok")
PY
timeout 280 compute-sanitizer --tool memcheck --print-limit 5 --log-file gpurun_out/r2_memcheck.log python /tmp/san.py > gpurun_out/r2_memcheck_run.log 2>&1; echo "memcheck rc=$?"
tail -2 gpurun_out/r2_memcheck_run.log; tail -4 gpurun_out/r2_memcheck.log
PPK_RALL_PERS=1 timeout 200 compute-sanitizer --tool memcheck --print-limit 5 --log-file gpurun_out/r2_memcheck_pers.log python /tmp/san.py > gpurun_out/r2_memcheck_pers_run.log 2>&1; echo "memcheck (persistent Riemann kernel) rc=$?"
tail -2 gpurun_out/r2_memcheck_pers_run.log; tail -3 gpurun_out/r2_memcheck_pers.log
timeout 280 compute-sanitizer --tool racecheck --print-limit 5 --log-file gpurun_out/r2_racecheck.log python /tmp/san.py > gpurun_out/r2_racecheck_run.log 2>&1; echo "racecheck rc=$?"
tail -2 gpurun_out/r2_racecheck_run.log; tail -6 gpurun_out/r2_racecheck.log
